"""GPU parity of two branches the reference pins with its own goldens (extracted into tests/golden/reference_screen_output.json):

* `Material averaging = project to Q1 only viscosity` (DGQ1 tables; matrix_free_operators.cc:306-325, local_smoothing.cc:99-111,
  :274-289): device tables against the oracle's on box meshes (2D / 3D), and tests/hollow_sphere_gmg_project_q1 on curved cells:
  35 iterations and the four printed error norms;
* `Enable prescribed dilation` (tests/prescribed_dilation_gmg): 35 / 36 iterations and the printed u_L2 / p_L2 errors, through
  smf_set_newton_data + smf_correct_stokes_rhs + volume normalisation.  The level operators must stay non-deviatoric
  (local_smoothing.cc:312) -- the iteration count is the witness (26 / 27 otherwise)."""
import numpy as np
import pytest

from oracle import assembled, coeff, mesh as omesh, postprocess, shell as oshell, solver
from aspect_b200 import SMF_ACTIVE, StokesMatrixFreeHandler, benchmarks, shell as pshell, synthetic
from util import golden, rel, seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,R", [(2, 3), (3, 2)])
@pytest.mark.parametrize("device_tables", [True, False])
def test_project_q1_tables_and_operators(dim, R, device_tables):
    hier, ohier = synthetic.BoxHierarchy(dim, R), omesh.Hierarchy(dim, R)
    xq = hier.active.quadrature_points()
    eta_q = np.exp(2.0 * np.sin(3.0 * xq[..., 0]) * np.cos(2.0 * xq[..., -1]) + xq[..., 1])
    eta_q[:: 7, 0] *= 30.0                                   # outliers: the projection overshoots and the clamp bites
    active, levels, emin, emax = coeff.viscosity_tables(ohier, eta_q, "project_q1")
    h = StokesMatrixFreeHandler(dim)
    h.device_cell_data = device_tables
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "project to Q1 only viscosity", 2.5)
    assert abs(h.minimum_viscosity / emin - 1) < 1e-14 and abs(h.maximum_viscosity / emax - 1) < 1e-14
    assert rel(h.get_cell_data(SMF_ACTIVE), active) < 1e-13
    for l in range(R + 1):
        assert h.get_cell_data(l).shape == levels[l].shape and rel(h.get_cell_data(l), levels[l]) < 1e-13
    assert (active == emax).any() or (active == emin).any()  # the clamp was active somewhere
    # operators with 3^dim viscosities per cell on every level
    h.build_preconditioner()
    ab = solver.AssembledBackend(ohier, active, levels, 2.5, "zero")
    for l in range(R + 1):
        xu, xp = seeded(hier.levels[l].n_u, l), seeded(hier.levels[l].n_p, 10 + l)
        du, dp = h.initialize_dof_vector(0, l), h.initialize_dof_vector(1, l)
        h.mg_matrices_A_block[l].vmult(du, h.initialize_dof_vector(0, l).from_host(xu))
        h.mg_matrices_Schur_complement[l].vmult(dp, h.initialize_dof_vector(1, l).from_host(xp))
        assert rel(du.to_host(), ab.apply_A(l, xu)) < 1e-12 and rel(dp.to_host(), ab.apply_M(l, xp)) < 1e-12
        assert rel(h.get_matrix_diagonal_inverse(l, 0), ab.dinv_A(l)) < 1e-12 and rel(h.get_matrix_diagonal_inverse(l, 1), ab.dinv_M(l)) < 1e-12
    h.close()


def test_hollow_sphere_project_q1_reference_golden():
    from oracle import hollow_sphere as hs
    g = golden("hollow_sphere_gmg_project_q1")
    m = 3
    ohier, hier = oshell.ShellHierarchy(1, 0.5, 1.0, n_coarse=12), pshell.ShellHierarchy(1, 0.5, 1.0, n_coarse=12)
    L = ohier.active
    eta_q, force_q, ps, u0 = benchmarks.hollow_sphere_problem(hier, m=m, averaging="project to Q1 only viscosity")
    active, levels, emin, emax = coeff.viscosity_tables(ohier, hs.material(L.xq, m)[0], "project_q1")
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "project to Q1 only viscosity", ps)
    assert rel(h.get_cell_data(SMF_ACTIVE), active) < 1e-12 and rel(h.get_cell_data(0), levels[0]) < 1e-12
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-8, "Number of cheap Stokes solver steps": 100,
                         "Maximum number of expensive Stokes solver steps": 0, "Pressure normalization": "surface"})
    r, b0, sol = h.initialize_dof_vector(), h.initialize_dof_vector(), h.initialize_dof_vector()
    h.assemble_stokes_rhs(force_q, r[0]); r[1].zero()
    b0[0].from_host(u0); b0[1].zero()
    h.correct_stokes_rhs(r, b0)
    sol[0].zero(); sol[1].zero()
    out = h.solve_device(sol, r)
    assert out["converged"] and abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][0][0]) <= 1     # 35
    u = sol[0].to_host() + u0
    h.set_pressure_normalization("surface")
    h.postprocess_pressure(sol[1], ps, True)
    for got, ref in zip(hs.errors(L, u, sol[1].to_host(), m), g["errors_uL1_pL1_uL2_pL2"][0]):
        assert abs(got / float(ref) - 1.0) < 2e-5
    h.close()


@pytest.mark.parametrize("R,mesh_index", [(3, 0), (4, 1)])
def test_prescribed_dilation_reference_golden(R, mesh_index):
    from oracle import prescribed_dilation as pd
    g = golden("prescribed_dilation_gmg")
    hier = synthetic.BoxHierarchy(2, R, [pd.EXTENT, pd.EXTENT])
    act = hier.active
    assert [act.n_u, act.n_p] == g["dofs_u_p"][mesh_index]
    oact = omesh.Hierarchy(2, R, extents=[pd.EXTENT, pd.EXTENT]).active
    assert np.array_equal(oact.cell_dofs_u, act.cell_dofs_u)
    eta_q, ps, rhs, u0 = pd.problem(oact)
    h = StokesMatrixFreeHandler(2)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    h.set_newton_data(enable_prescribed_dilation=True, dilation_lhs_term_table=np.zeros((act.n_cells, 9)))
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-12, "Number of cheap Stokes solver steps": 200,
                         "Maximum number of expensive Stokes solver steps": 0})
    r, b0, sol = h.initialize_dof_vector(), h.initialize_dof_vector(), h.initialize_dof_vector()
    rhs[act.constrained_u("zero").astype(np.int64)] = 0.0        # constrained rows of the assembled right-hand side stay empty
    r[0].from_host(rhs[:act.n_u]); r[1].from_host(rhs[act.n_u:])
    b0[0].from_host(u0); b0[1].zero()
    h.correct_stokes_rhs(r, b0)
    sol[0].zero(); sol[1].zero()
    out = h.solve_device(sol, r)
    assert out["converged"]
    assert out["iterations"] == g["stokes_iterations_cheap_expensive"][2 * mesh_index][0], out["iterations"]     # 35 / 36
    u = sol[0].to_host() + u0
    h.set_pressure_normalization("volume")
    h.postprocess_pressure(sol[1], ps, True)
    eu, ep = pd.errors(oact, u, sol[1].to_host())
    gu, gp = (float(v) for v in g["errors_uL2_pL2"][mesh_index])
    assert abs(eu / gu - 1) < 1e-5 and abs(ep / gp - 1) < 1e-5, (eu, ep, gu, gp)
    h.close()
