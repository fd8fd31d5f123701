"""Committed golden vectors (tests/golden/operators_*.npz, made by tools/make_operator_goldens.py from the oracle's
assembled sparse operators): the C matrix-free oracle on the CPU, and the CUDA path through the C ABI on the GPU,
must both reproduce them.  Tolerance: relative 1e-12 in double (BASELINE.json north_star: "a single operator
application matches ... to relative 1e-12")."""
import glob
import os

import numpy as np
import pytest

from oracle import coeff, cpu, mesh as omesh
from util import rel

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operators_*.npz")))
TOL = 1e-12
AVG = {"harmonic": "harmonic average", "none": "none"}


def _load(path):
    g = np.load(path)
    return g, int(g["dim"]), int(g["R"]), str(g["bc"]), str(g["averaging"]), float(g["pressure_scaling"])


def test_fixtures_exist():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
def test_oracle_matrix_free_reproduces_golden_vectors(path):
    g, dim, R, bc, avg, ps = _load(path)
    hier = omesh.Hierarchy(dim, R)
    active, levels, _, _ = coeff.viscosity_tables(hier, g["eta_q"], avg)
    ob = cpu.backend_from_hierarchy(hier, active, levels, ps, bc)
    yu, yp = ob.stokes(g["xu"], g["xp"])
    assert rel(yu, g["stokes_u"]) < TOL and rel(yp, g["stokes_p"]) < TOL
    assert rel(ob.A_active(g["xu"]), g["A_active"]) < TOL
    assert rel(ob.BT(g["xp"]), g["BT"]) < TOL
    assert rel(ob.M_active(g["xp"]), g["mass_active"]) < TOL
    for l in range(R + 1):
        assert rel(ob.apply_A(l, g[f"level{l}_xu"]), g[f"level{l}_A"]) < TOL
        assert rel(ob.apply_M(l, g[f"level{l}_xp"]), g[f"level{l}_mass"]) < TOL
        assert rel(ob.dinv_A(l), g[f"level{l}_dinv_A"]) < TOL and rel(ob.dinv_M(l), g[f"level{l}_dinv_M"]) < TOL
        if l > 0:
            assert rel(ob.prolongate_u(l, g[f"level{l}_coarse_u"]), g[f"level{l}_prolongate_u"]) < TOL
            assert rel(ob.prolongate_p(l, g[f"level{l}_coarse_p"]), g[f"level{l}_prolongate_p"]) < TOL
            assert rel(ob.restrict_u(l, g[f"level{l}_xu"]), g[f"level{l}_restrict_u"]) < TOL
            assert rel(ob.restrict_p(l, g[f"level{l}_xp"]), g[f"level{l}_restrict_p"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("structured", [True, False], ids=["structured_bricks", "general_batches"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[10:-4] for p in GOLDEN])
def test_cuda_path_reproduces_golden_vectors(path, structured):
    from aspect_b200 import SMF_ACTIVE, StokesMatrixFreeHandler, synthetic
    g, dim, R, bc, avg, ps = _load(path)
    hier = synthetic.BoxHierarchy(dim, R)
    h = StokesMatrixFreeHandler(dim)
    h.allow_structured_bricks = structured
    h.setup_dofs(hier, bc)
    h.assemble(g["eta_q"], AVG[avg], ps)
    h.build_preconditioner()
    src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
    src[0].from_host(g["xu"]); src[1].from_host(g["xp"])
    h.stokes_matrix.vmult(dst, src)
    assert rel(dst[0].to_host(), g["stokes_u"]) < TOL and rel(dst[1].to_host(), g["stokes_p"]) < TOL
    h.A_block_matrix.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), g["A_active"]) < TOL
    h.BT_block.vmult(dst, src)
    assert rel(dst[0].to_host(), g["BT"]) < TOL
    h.Schur_complement_block_matrix.vmult(dst[1], src[1])
    assert rel(dst[1].to_host(), g["mass_active"]) < TOL
    for l in range(R + 1):
        xu, xp = h.initialize_dof_vector(0, l).from_host(g[f"level{l}_xu"]), h.initialize_dof_vector(1, l).from_host(g[f"level{l}_xp"])
        yu, yp = h.initialize_dof_vector(0, l), h.initialize_dof_vector(1, l)
        h.mg_matrices_A_block[l].vmult(yu, xu)
        h.mg_matrices_Schur_complement[l].vmult(yp, xp)
        assert rel(yu.to_host(), g[f"level{l}_A"]) < TOL and rel(yp.to_host(), g[f"level{l}_mass"]) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 0), g[f"level{l}_dinv_A"]) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 1), g[f"level{l}_dinv_M"]) < TOL
        if l > 0:
            cu = h.initialize_dof_vector(0, l - 1).from_host(g[f"level{l}_coarse_u"])
            cp = h.initialize_dof_vector(1, l - 1).from_host(g[f"level{l}_coarse_p"])
            h.prolongate(0, l, yu, cu); h.prolongate(1, l, yp, cp)
            assert rel(yu.to_host(), g[f"level{l}_prolongate_u"]) < TOL and rel(yp.to_host(), g[f"level{l}_prolongate_p"]) < TOL
            cu.zero(); cp.zero()
            h.restrict_and_add(0, l, cu, xu); h.restrict_and_add(1, l, cp, xp)
            assert rel(cu.to_host(), g[f"level{l}_restrict_u"]) < TOL and rel(cp.to_host(), g[f"level{l}_restrict_p"]) < TOL
    h.close()


# ---- curved cells: tests/golden/curved_*.npz (spherical shell of tests/hollow_sphere_gmg, restated SphericalManifold /
# MappingQ(4) geometry; made by tools/make_operator_goldens.py from the oracle's assembled curved-geometry operators) -----
CURVED = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "curved_*.npz")))


def _shell(g):
    from aspect_b200 import shell
    return shell.ShellHierarchy(int(g["R"]), float(g["inner_radius"]), float(g["outer_radius"]), n_coarse=int(g["n_coarse"]))


@pytest.mark.parametrize("path", CURVED, ids=[os.path.basename(p)[7:-4] for p in CURVED])
def test_shell_driver_reproduces_the_fixture_geometry(path):
    """The vectorised shell driver (what the GPU tests and the bench feed to the C ABI) against the committed fixture:
    bit-exact cell->DoF map, inverse Jacobians and JxW of the pinned geometry."""
    g = np.load(path)
    act = _shell(g).active
    assert np.array_equal(act.cell_dofs_u, g["cell_dofs_u"])
    assert np.abs(act.jinv - g["jinv_active"]).max() < 1e-11 and np.abs(act.jxw - g["jxw_active"]).max() < 1e-14


def test_curved_fixtures_exist():
    assert len(CURVED) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [True, False], ids=["sum_factorised", "general_kernel"])
@pytest.mark.parametrize("path", CURVED, ids=[os.path.basename(p)[7:-4] for p in CURVED])
def test_cuda_path_reproduces_curved_golden_vectors(path, fast):
    from aspect_b200 import SMF_ACTIVE, StokesMatrixFreeHandler
    g = np.load(path)
    hier = _shell(g)
    R, avg, ps = int(g["R"]), str(g["averaging"]), float(g["pressure_scaling"])
    h = StokesMatrixFreeHandler(3)
    h.curved_sum_factorization = fast
    h.setup_dofs(hier, "zero")
    h.assemble(g["eta_q"], AVG[avg], ps)
    h.build_preconditioner()
    assert rel(h.get_cell_data(SMF_ACTIVE), g["active_table"]) < 1e-14
    assert abs(h.minimum_viscosity / g["eta_min_max"][0] - 1) < 1e-14 and abs(h.maximum_viscosity / g["eta_min_max"][1] - 1) < 1e-14
    src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
    src[0].from_host(g["xu"]); src[1].from_host(g["xp"])
    h.stokes_matrix.vmult(dst, src)
    assert rel(dst[0].to_host(), g["stokes_u"]) < TOL and rel(dst[1].to_host(), g["stokes_p"]) < TOL
    h.A_block_matrix.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), g["A_active"]) < TOL
    h.BT_block.vmult(dst, src)
    assert rel(dst[0].to_host(), g["BT"]) < TOL
    h.Schur_complement_block_matrix.vmult(dst[1], src[1])
    assert rel(dst[1].to_host(), g["mass_active"]) < TOL
    for l in range(R + 1):
        xu, xp = h.initialize_dof_vector(0, l).from_host(g[f"level{l}_xu"]), h.initialize_dof_vector(1, l).from_host(g[f"level{l}_xp"])
        yu, yp = h.initialize_dof_vector(0, l), h.initialize_dof_vector(1, l)
        h.mg_matrices_A_block[l].vmult(yu, xu)
        h.mg_matrices_Schur_complement[l].vmult(yp, xp)
        assert rel(yu.to_host(), g[f"level{l}_A"]) < TOL and rel(yp.to_host(), g[f"level{l}_mass"]) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 0), g[f"level{l}_dinv_A"]) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 1), g[f"level{l}_dinv_M"]) < TOL
        if l > 0:
            cu = h.initialize_dof_vector(0, l - 1).from_host(g[f"level{l}_coarse_u"])
            cp = h.initialize_dof_vector(1, l - 1).from_host(g[f"level{l}_coarse_p"])
            h.prolongate(0, l, yu, cu); h.prolongate(1, l, yp, cp)
            assert rel(yu.to_host(), g[f"level{l}_prolongate_u"]) < TOL and rel(yp.to_host(), g[f"level{l}_prolongate_p"]) < TOL
            cu.zero(); cp.zero()
            h.restrict_and_add(0, l, cu, xu); h.restrict_and_add(1, l, cp, xp)
            assert rel(cu.to_host(), g[f"level{l}_restrict_u"]) < TOL and rel(cp.to_host(), g[f"level{l}_restrict_p"]) < TOL
    h.close()
