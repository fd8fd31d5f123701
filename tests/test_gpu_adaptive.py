"""GPU parity on adaptively refined meshes (SURVEY 8(a) a12 + hanging-node constraints of the north star) against
oracle/adaptive.py, which fixes the semantics (MatrixFreeOperators::Base: A_mf = C^T A C + I_c; local smoothing with
refinement-edge rows and interface operators; MGTransfer copy indices):

  * active-level operators (A, Stokes, B^T, mass) with hanging nodes, 1e-12;
  * level operators A_l = (I - E) A_bc (I - E) + E, their inverse diagonals, partial-level transfers;
  * both adaptive V-cycles and the block preconditioner, 1e-10;
  * the complete solve: same outer iterations as the oracle, solution 1e-8.
Meshes: 3D box, refinement 1 + 2 adaptive rounds around a corner (3 levels of cells in one mesh) and refinement 2 + 1 round
around an nsinker centre (the mesh of tests/test_adaptive_oracle.py); 2D box with a refined corner / a refined band (the active
level goes in with per-cell geometry, i.e. through the 2D instantiations of the general-geometry kernels)."""
import numpy as np
import pytest

from oracle import adaptive as oad, coeff
from aspect_b200 import SMF_ACTIVE, StokesMatrixFreeHandler, adaptive as pad
from util import rel, seeded

pytestmark = pytest.mark.gpu

MESHES = {"corner": (3, 1, 2, lambda x, l: np.linalg.norm(x) < 0.6),
          "sinker": (3, 2, 1, lambda x, l: np.linalg.norm(x - coeff.NSINKER_CENTERS[4]) < 0.3),
          "corner2d": (2, 2, 2, lambda x, l: np.linalg.norm(x) < 0.6),
          "band2d": (2, 3, 1, lambda x, l: abs(x[0] - 0.5) < 0.2 and x[1] > 0.4)}


def _setup(name, bc="zero", ps=1.7):
    dim, base, rounds, marker = MESHES[name]
    box = oad.AdaptiveBox.refined(dim, base, marker, rounds)
    hier = pad.AdaptiveBoxHierarchy.refined(dim, base, marker, rounds)
    assert np.array_equal(hier.active.cell_dofs_u, box.cell_dofs_u) and len(box.hanging_u) > 0
    eta = 10.0 ** np.random.default_rng(3).uniform(-1.0, 1.0, box.n_cells)
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(np.repeat(eta[:, None], 3 ** dim, axis=1), "harmonic average", ps)
    return box, hier, eta, h


@pytest.mark.parametrize("name", ["corner", "sinker", "corner2d", "band2d"])
@pytest.mark.parametrize("bc", ["zero", "tangential"])
def test_active_operators_with_hanging_nodes(name, bc):
    box, hier, eta, h = _setup(name, bc)
    ops = box.assemble(eta, 1.7, bc)
    nu, npr = box.n_u, box.n_p
    xu, xp = seeded(nu, 1), seeded(npr, 2)                        # constrained entries carry values: dst = src there
    s, d = h.initialize_dof_vector(), h.initialize_dof_vector()
    s[0].from_host(xu); s[1].from_host(xp)
    h.A_block_matrix.vmult(d[0], s[0])
    assert rel(d[0].to_host(), ops["A"] @ xu) < 1e-12
    assert np.array_equal(s[0].to_host(), xu)                     # the source vector comes back untouched
    h.Schur_complement_block_matrix.vmult(d[1], s[1])
    assert rel(d[1].to_host(), ops["M"] @ xp) < 1e-12
    h.stokes_matrix.vmult(d, s)
    want = ops["K"] @ np.concatenate([xu, xp])
    assert rel(d[0].to_host(), want[:nu]) < 1e-12 and rel(d[1].to_host(), want[nu:]) < 1e-12
    assert np.array_equal(s[0].to_host(), xu) and np.array_equal(s[1].to_host(), xp)
    h.BT_block.vmult(d, s)
    assert rel(d[0].to_host(), ops["BT"] @ xp) < 1e-12
    h.close()


@pytest.mark.parametrize("name", ["corner", "sinker", "corner2d", "band2d"])
def test_level_operators_diagonals_and_transfers(name):
    box, hier, eta, h = _setup(name)
    oh = oad.AdaptiveHierarchy(box, "zero")
    lops = oh.level_operators(oh.level_viscosity(eta), 1.7)
    h.build_preconditioner()
    for l, (O, L) in enumerate(zip(lops, hier.levels)):
        xu, xp = seeded(L.n_u, 10 + l), seeded(L.n_p, 20 + l)
        du, dp = h.initialize_dof_vector(0, l), h.initialize_dof_vector(1, l)
        su, sp = h.initialize_dof_vector(0, l).from_host(xu), h.initialize_dof_vector(1, l).from_host(xp)
        h.mg_matrices_A_block[l].vmult(du, su)
        h.mg_matrices_Schur_complement[l].vmult(dp, sp)
        assert rel(du.to_host(), O["A"] @ xu) < 1e-12 and rel(dp.to_host(), O["M"] @ xp) < 1e-12
        assert np.array_equal(su.to_host(), xu) and np.array_equal(sp.to_host(), xp)
        assert rel(h.get_matrix_diagonal_inverse(l, 0), 1.0 / O["A"].diagonal()) < 1e-12
        assert rel(h.get_matrix_diagonal_inverse(l, 1), 1.0 / O["M"].diagonal()) < 1e-12
        if l > 0:
            C = hier.levels[l - 1]
            for which, fam, nf, nc in ((0, "u", L.n_u, C.n_u), (1, "p", L.n_p, C.n_p)):
                P = oh.prolongation(l, fam)
                xc, xf = seeded(nc, 30 + l), seeded(nf, 40 + l)
                f, c = h.initialize_dof_vector(which, l), h.initialize_dof_vector(which, l - 1)
                f.zero()
                h.prolongate(which, l, f, h.initialize_dof_vector(which, l - 1).from_host(xc))
                assert rel(f.to_host(), P @ xc) < 1e-13
                c.zero()
                h.restrict_and_add(which, l, c, h.initialize_dof_vector(which, l).from_host(xf))
                assert rel(c.to_host(), P.T @ xf) < 1e-13
    h.close()


@pytest.mark.parametrize("name", ["corner", "sinker", "corner2d", "band2d"])
def test_adaptive_vcycles_and_preconditioner(name):
    box, hier, eta, h = _setup(name)
    oh = oad.AdaptiveHierarchy(box, "zero")
    lops = oh.level_operators(oh.level_viscosity(eta), 1.7)
    mgA, mgS = oad.AdaptiveVCycle(oh, lops, "u", 1), oad.AdaptiveVCycle(oh, lops, "p", 4)
    h.build_preconditioner()
    h.setup_smoothers()
    for l in range(len(hier.levels)):
        for which, mg in ((0, mgA), (1, mgS)):
            lam = h.eigenvalue_estimates(l, which)
            assert abs(lam[1] / (1.2 * mg.smoothers[l].max_eig_est) - 1) < 1e-9 and int(lam[2]) == mg.smoothers[l].cg_iterations
    ops = box.assemble(eta, 1.7, "zero")
    ru, rp = seeded(box.n_u, 5), seeded(box.n_p, 6)
    ru[ops["constrained_u"]] = 0.0; rp[ops["constrained_p"]] = 0.0
    du, dp = h.initialize_dof_vector()
    h.prec_A.vmult(du, h.initialize_dof_vector(0).from_host(ru))
    h.prec_Schur.vmult(dp, h.initialize_dof_vector(1).from_host(rp))
    assert rel(du.to_host(), mgA.vmult(ru)) < 1e-10
    assert rel(dp.to_host(), mgS.vmult(rp)) < 1e-10
    h.close()


@pytest.mark.parametrize("name", ["sinker", "band2d"])
def test_adaptive_solve_matches_oracle(name):
    """3D: the mesh of tests/test_adaptive_oracle.py::test_adaptive_solve_converges_to_the_direct_solution (106 cells, refinement
    2 + one adaptive round); 2D: refinement 3 + one round in a band.  Viscosity contrast 100 from cell to cell."""
    box, hier, eta, h = _setup(name, ps=1.0)
    ops = box.assemble(eta, 1.0, "zero")
    rhs = np.concatenate([ops["Cu"].T @ np.sin(np.arange(box.n_u)), np.zeros(box.n_p)])
    rhs[ops["constrained_u"]] = 0.0
    # Solver tolerance 1e-10: two runs that stop at 1e-8 agree to ~1e-8 only if they follow the same Krylov path to the last bit,
    # and the level kernels of this path accumulate with unordered REDs (observed: 0.7e-8 .. 1.1e-8 from run to run); two decades
    # below the 1e-8 criterion of the comparison it is the discretised problem that is compared, not the stopping point.
    want = oad.block_gmg_solve(box, eta, rhs, "zero", 1.0, 1e-10)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-10, "Number of cheap Stokes solver steps": 200,
                         "Maximum number of expensive Stokes solver steps": 0})
    sol, out = h.solve(rhs)
    assert out["converged"] and want["converged"]
    assert abs(out["iterations"] - want["iterations"]) <= 1, (out["iterations"], want["iterations"])
    nu = box.n_u
    free = np.setdiff1d(np.arange(nu), ops["constrained_u"])
    e_free = rel(sol[:nu][free], want["x"][:nu][free])
    assert e_free < 1e-8, e_free
    # distribute(): the hanging entries of the returned solution carry the interpolated values
    full = ops["Cu"] @ want["x"][:nu]
    e_full = rel(sol[:nu], full)
    assert e_full < 1e-8, (e_full, e_free)
    h.close()


@pytest.mark.parametrize("which", [2, 3])
def test_muparser_temperature_example_reference_golden(which):
    """The reference's adaptive golden tests/muparser_temperature_example (second and third solve: 260 cells on 5 levels, 920 on 6;
    5 x 1 coarse cells; meshes given by a `minimum refinement function`, see oracle/muparser_example.py) through the C ABI:
    local smoothing with hanging nodes, refinement edges, interface operators and copy indices on the device.  Golden: 9 and 9
    GMRES iterations; the oracle takes 9 and 10; the device must repeat the oracle (+-1) and stay within 2 of the golden."""
    from oracle import muparser_example as mx
    from aspect_b200 import benchmarks
    from util import golden
    gold = golden("muparser_temperature_example")["stokes_iterations_cheap_expensive"][which - 1][0]
    box, want = mx.solve(which)
    hier = pad.AdaptiveBoxHierarchy(2, benchmarks.muparser_example_leaves(which), **benchmarks.MUPARSER_BOX)
    act = hier.active
    assert np.array_equal(act.cell_dofs_u, box.cell_dofs_u) and len(act.hanging_u) == len(box.hanging_u) > 0
    eta_q, rhs, ps = benchmarks.muparser_example_problem(hier)
    h = StokesMatrixFreeHandler(2)
    h.setup_dofs(hier, "tangential")
    h.assemble(eta_q, "harmonic average", ps)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-6, "Number of cheap Stokes solver steps": 200,
                         "Maximum number of expensive Stokes solver steps": 0, "GMRES solver restart length": 100})
    rhs = h.condense_rhs(rhs)
    rhs[act.constrained_u("tangential").astype(np.int64)] = 0.0
    sol, out = h.solve(rhs)
    print(f"muparser_temperature_example solve {which}: device {out['iterations']}, oracle {want['iterations']}, reference {gold}")
    assert out["converged"] and want["converged"]
    assert abs(out["iterations"] - want["iterations"]) <= 1 and abs(out["iterations"] - gold) <= 2, (out["iterations"], want["iterations"], gold)
    nu = box.n_u
    full = want["ops"]["Cu"] @ want["x"][:nu]
    assert rel(sol[:nu], full) < 1e-4, rel(sol[:nu], full)            # (both stop at a relative 1e-6 of a mostly hydrostatic load)
    h.close()
