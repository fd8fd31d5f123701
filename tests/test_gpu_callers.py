"""GPU parity of the callers either side of the solve: coefficient-table pipeline, right-hand-side
assembly and pressure post-processing on the device, against the oracle's numpy restatements."""
import numpy as np
import pytest

from oracle import assembled, coeff, mesh as omesh, postprocess, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, cell_data, synthetic
from aspect_b200._lib import SMF_ACTIVE
from util import rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,R", [(2, 4), (3, 3)])
@pytest.mark.parametrize("averaging", ["none", "arithmetic average", "harmonic average", "geometric average", "harmonic average only viscosity"])
def test_fill_cell_data_on_device(dim, R, averaging):
    """smf_fill_cell_data = material averaging (interface.cc:523-650) + DGQ0 projection (utilities.cc:2738-2812) + level
    transfer (local_smoothing.cc:211-294) + min / max (matrix_free_operators.cc:248-262)."""
    hier, ohier = synthetic.BoxHierarchy(dim, R), omesh.Hierarchy(dim, R)
    rng = np.random.default_rng(2)
    eta_q = np.exp(rng.uniform(-4, 4, size=(hier.active.n_cells, 3 ** dim)))
    if averaging.startswith("harmonic"):
        eta_q[3, 5] = 0.0                                    # "if one of the values is zero, the average is 0.0"
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, averaging, 1.0)
    op = averaging.replace(" only viscosity", "").replace(" average", "")
    active, levels, emin, emax = coeff.viscosity_tables(ohier, eta_q, op)
    assert rel(h.get_cell_data(SMF_ACTIVE), active) < 1e-14
    for l in range(R + 1):
        assert rel(h.get_cell_data(l), levels[l]) < 1e-14
    assert abs(h.minimum_viscosity - emin) <= 1e-15 * abs(emin) and abs(h.maximum_viscosity - emax) <= 1e-14 * abs(emax)
    # the host pipeline of the plugin layer gives the same tables
    children = [None] + [hier.children(l) for l in range(1, R + 1)]
    a2, l2, mn2, mx2 = cell_data.fill_cell_data(dim, children, eta_q, averaging)
    assert rel(h.get_cell_data(SMF_ACTIVE), a2) < 1e-14 and rel(h.get_cell_data(0), l2[0]) < 1e-14
    h.close()


@pytest.mark.parametrize("dim,R,bc", [(2, 4, "tangential"), (3, 3, "zero")])
def test_assemble_rhs_on_device(dim, R, bc):
    hier, ohier = synthetic.BoxHierarchy(dim, R), omesh.Hierarchy(dim, R)
    act = ohier.active
    rng = np.random.default_rng(7)
    force = rng.uniform(-1, 1, size=(act.n_cells, 3 ** dim, dim))
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    f = h.initialize_dof_vector(0)
    h.assemble_stokes_rhs(force, f)
    want = np.zeros(act.n_u)
    for c in range(dim):
        g = np.zeros(dim); g[c] = 1.0
        want += assembled.rhs_gravity(act, force[:, :, c], g)
    want[act.constrained_u(bc)] = 0.0
    assert rel(f.to_host(), want) < 1e-13
    h.close()


def test_solve_with_device_rhs_and_surface_normalization():
    """SolCx R=4 end to end on the device side of the boundary: right-hand side from the force density, solve, pressure
    back in physical units and `Pressure normalization = surface`; errors as printed by tests/sol_cx_4_gmg."""
    dim, R, bc, avg = 2, 4, "tangential", "harmonic average only viscosity"
    hier, ohier = synthetic.BoxHierarchy(dim, R), omesh.Hierarchy(dim, R)
    act = ohier.active
    eta_q, rho_q = benchmarks.solcx_material(hier.active.quadrature_points())
    ps = benchmarks.pressure_scaling_box(hier.active, benchmarks.solcx_material)
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, avg, ps)
    h.build_preconditioner()
    force = np.zeros((act.n_cells, 3 ** dim, dim)); force[:, :, 1] = -rho_q
    fu = h.initialize_dof_vector(0)
    h.assemble_stokes_rhs(force, fu)
    rhs = np.concatenate([fu.to_host(), np.zeros(act.n_p)])
    _, want_rhs, _ = benchmarks.solcx_problem(hier, averaging=avg)
    assert rel(rhs, want_rhs) < 1e-13
    h.parameters.update({"Linear solver tolerance": 1e-7, "Pressure normalization": "surface"})
    sol, out = h.solve(rhs)
    assert out["converged"] and out["iterations"] == 15
    ob = osolver.AssembledBackend(ohier, *coeff.viscosity_tables(ohier, eta_q, "harmonic")[:2], ps, bc)
    oout = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-7, restart=100, max_its=1000)
    po, mean = postprocess.normalize_pressure_surface(act, oout["x"][act.n_u:] * ps)
    assert rel(sol[act.n_u:], po) < 1e-8
    assert abs(out["pressure_normalization_adjustment"] + mean) <= 1e-8 * max(abs(mean), 1e-300)
    err = postprocess.solcx_errors(act, sol[:act.n_u], sol[act.n_u:])
    assert [f"{e:.6e}" for e in err] == ["1.125999e-06", "1.101883e-01", "1.670011e-06", "1.106214e-01"]
    h.close()


def test_volume_normalization():
    dim, R = 3, 2
    hier, ohier = synthetic.BoxHierarchy(dim, R), omesh.Hierarchy(dim, R)
    act = ohier.active
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, "zero")
    p = np.random.default_rng(3).uniform(-1, 1, act.n_p)
    v = h.initialize_dof_vector(1)
    v.from_host(p)
    h.set_pressure_normalization("volume")
    adj = h.postprocess_pressure(v, pressure_scaling=2.5, normalize=True)
    want, mean = postprocess.normalize_pressure_volume(act, 2.5 * p)
    assert rel(v.to_host(), want) < 1e-13 and abs(adj + mean) < 1e-13
    h.close()
