"""GPU parity of the general cell kernel: compressible, prescribed-dilation and Newton-derivative branches of
StokesOperator::local_apply (matrix_free_operators.cc:376-547) and the deviatoric A block (:974-1009), against the
oracle's restatement (oracle/newton.py, itself anchored on the assembled operator and on symmetry)."""
import numpy as np
import pytest

from oracle import assembled, coeff, cpu, mesh as omesh, newton, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from test_newton_oracle import random_newton_data
from util import rel, seeded

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _setup(dim, R, bc, averaging, compressible, seed=1):
    hier = synthetic.BoxHierarchy(dim, R)
    ohier = omesh.Hierarchy(dim, R)
    rng = np.random.default_rng(seed)
    eta_q = np.exp(rng.uniform(-2, 2, size=(hier.active.n_cells, 3 ** dim)))
    ps = 1.3
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, averaging, ps, is_compressible=compressible)
    op = {"none": "none", "harmonic average": "harmonic"}[averaging]
    eta_active, eta_levels, _, _ = coeff.viscosity_tables(ohier, eta_q, op)
    return hier, ohier, h, eta_active, eta_levels, ps


def _vectors(h, act, bc, seed):
    xu, xp = seeded(act.n_u, seed), seeded(act.n_p, seed + 1)
    src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
    src[0].from_host(xu); src[1].from_host(xp)
    return xu, xp, src, dst


@pytest.mark.parametrize("dim,R,bc", [(2, 3, "tangential"), (3, 2, "zero")])
@pytest.mark.parametrize("averaging", ["none", "harmonic average"])
def test_compressible_operators(dim, R, bc, averaging):
    hier, ohier, h, eta_active, eta_levels, ps = _setup(dim, R, bc, averaging, True)
    h.build_preconditioner()
    act = ohier.active
    cons = act.constrained_u(bc)
    xu, xp, src, dst = _vectors(h, act, bc, 3)
    h.stokes_matrix.vmult(dst, src)
    yu, yp = newton.stokes_apply(act, eta_active, ps, cons, xu, xp, True)
    assert rel(np.concatenate([dst[0].to_host(), dst[1].to_host()]), np.concatenate([yu, yp])) < TOL
    h.A_block_matrix.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), newton.A_apply(act, eta_active, cons, xu, True)) < TOL
    for l, lev in enumerate(ohier.levels):            # level operators and their diagonals (deviatoric)
        cl = lev.constrained_u(bc)
        x = seeded(lev.n_u, 10 + l)
        s, d = h.initialize_dof_vector(0, l), h.initialize_dof_vector(0, l)
        s.from_host(x)
        h.mg_matrices_A_block[l].vmult(d, s)
        assert rel(d.to_host(), newton.A_apply(lev, eta_levels[l], cl, x, True)) < TOL
        A = assembled.assemble_A(lev, eta_levels[l], cl, True)
        assert rel(1.0 / h.get_matrix_diagonal_inverse(l, 0), A.diagonal()) < TOL
    h.close()


FLAGS = [dict(enable_newton_derivatives=True, symmetrize_newton_system=s, average_newton_factors=a, enable_prescribed_dilation=d)
         for s in (False, True) for a in (False, True) for d in (False, True)] + [dict(enable_prescribed_dilation=True)]


@pytest.mark.parametrize("dim,R,bc", [(2, 3, "tangential"), (3, 2, "zero")])
@pytest.mark.parametrize("compressible", [False, True])
@pytest.mark.parametrize("flags", FLAGS, ids=lambda f: "".join(k[7] if v else "-" for k, v in sorted(f.items())))
def test_newton_and_dilation_branches(dim, R, bc, compressible, flags):
    hier, ohier, h, eta_active, eta_levels, ps = _setup(dim, R, bc, "none", compressible, seed=4)
    act = ohier.active
    cons = act.constrained_u(bc)
    nd = random_newton_data(act, np.random.default_rng(9), **flags)
    h.set_newton_data(**{k: getattr(nd, k) for k in (
        "enable_newton_derivatives", "symmetrize_newton_system", "average_newton_factors", "enable_prescribed_dilation",
        "strain_rate_table", "newton_factor_wrt_strain_rate_table", "newton_factor_wrt_pressure_table", "dilation_lhs_term_table",
        "dilation_derivative_wrt_strain_rate_table", "dilation_derivative_wrt_pressure_table")})
    xu, xp, src, dst = _vectors(h, act, bc, 5)
    h.stokes_matrix.vmult(dst, src)
    yu, yp = newton.stokes_apply(act, eta_active, ps, cons, xu, xp, compressible, nd)
    assert rel(np.concatenate([dst[0].to_host(), dst[1].to_host()]), np.concatenate([yu, yp])) < TOL
    # the A block only sees the deviatoric switch (matrix_free_operators.cc:999-1005)
    h.A_block_matrix.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), newton.A_apply(act, eta_active, cons, xu, compressible, nd.enable_prescribed_dilation)) < TOL
    # correct_stokes_rhs reads u0 plainly and adds nothing to constrained rows
    r = h.initialize_dof_vector()
    r[0].zero(); r[1].zero()
    h.correct_stokes_rhs(r, src)
    cu, cp = newton.stokes_apply(act, eta_active, ps, cons, xu, xp, compressible, nd, plain=True)
    assert rel(np.concatenate([r[0].to_host(), r[1].to_host()]), -np.concatenate([cu, cp])) < TOL
    # clearing the tables restores the plain operator
    h.set_newton_data()
    h.stokes_matrix.vmult(dst, src)
    yu, yp = newton.stokes_apply(act, eta_active, ps, cons, xu, xp, compressible)
    assert rel(np.concatenate([dst[0].to_host(), dst[1].to_host()]), np.concatenate([yu, yp])) < TOL
    h.close()


def test_compressible_solve_matches_oracle():
    """A complete block-GMG solve with is_compressible = true (deviatoric A block on every level, general kernel)."""
    dim, R, bc, avg = 2, 3, "tangential", "harmonic average"
    hier = synthetic.BoxHierarchy(dim, R)
    eta_q, rhs, ps = benchmarks.solcx_problem(hier, averaging=avg)
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, avg, ps, is_compressible=True)
    h.build_preconditioner()
    h.parameters["Linear solver tolerance"] = 1e-7
    sol, out = h.solve(rhs)
    ohier = omesh.Hierarchy(dim, R)
    active, levels, _, _ = coeff.viscosity_tables(ohier, eta_q, "harmonic")
    ob = osolver.AssembledBackend(ohier, active, levels, ps, bc, compressible=True)
    oout = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-7, restart=100, max_its=1000)
    assert out["converged"] and out["iterations"] == oout["iterations"]
    assert rel(sol[:hier.active.n_u], oout["x"][:hier.active.n_u]) < 1e-8
    h.close()


@pytest.mark.parametrize("dim,R", [(2, 3), (3, 2)])
def test_free_surface_face_term_and_bicgstab(dim, R):
    """StokesOperator::local_apply_boundary_face (matrix_free_operators.cc:563-602) on the top faces of the box against the
    oracle's restatement, and the expensive phase with BiCGStab for the A block (block_stokes_preconditioner.h:119-132):
    same FGMRES iterations and inner-iteration totals as the oracle."""
    hier = synthetic.BoxHierarchy(dim, R)
    ohier = omesh.Hierarchy(dim, R)
    act, L = hier.active, ohier.active
    rng = np.random.default_rng(7)
    eta_q = 10.0 ** rng.uniform(-1, 1, (act.n_cells, 3 ** dim))
    bc = "tangential"
    cons = L.constrained_u(bc)
    cells = np.flatnonzero(act.cells[:, dim - 1] == (1 << R) - 1)                 # cells under the top boundary
    face_no = 2 * (dim - 1) + 1
    tab = rng.uniform(-1.0, 1.0, (len(cells), 3 ** (dim - 1), dim)) * 0.05
    tab[..., dim - 1] -= 0.2                                                       # g ~ rho dt theta |g| g_hat, pointing down
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, "harmonic average", 1.3)
    h.set_free_surface_faces(cells, np.full(len(cells), face_no), tab)
    active, levels, _, _ = coeff.viscosity_tables(ohier, eta_q, "harmonic")
    xu, xp = seeded(act.n_u, 3), seeded(act.n_p, 4)
    s, d = h.initialize_dof_vector(), h.initialize_dof_vector()
    s[0].from_host(xu); s[1].from_host(xp)
    h.stokes_matrix.vmult(d, s)
    ob = cpu.backend_from_hierarchy(ohier, active, levels, 1.3, bc)
    xr = xu.copy(); xr[cons] = 0.0
    yu, yp = ob.stokes(xr, xp)
    yu[cons] = xu[cons]
    yu = newton.free_surface_apply_add(L, cells, face_no, tab, cons, xu, yu)
    assert rel(d[0].to_host(), yu) < 1e-12 and rel(d[1].to_host(), yp) < 1e-12
    # without faces the term is gone again
    h.set_free_surface_faces([], [], np.zeros((0, 3 ** (dim - 1), dim)))
    h.stokes_matrix.vmult(d, s)
    yu0, _ = ob.stokes(xr, xp); yu0[cons] = xu[cons]
    assert rel(d[0].to_host(), yu0) < 1e-12
    # expensive phase with BiCGStab on the A block (forced: the operator itself is the symmetric one)
    h.build_preconditioner()
    h.force_nonsymmetric_A_block_solver(True)
    h.parameters.update({"Linear solver tolerance": 1e-7, "Number of cheap Stokes solver steps": 0,
                         "Maximum number of expensive Stokes solver steps": 100})
    rhs = np.concatenate([seeded(act.n_u, 8), np.zeros(act.n_p)])
    rhs[cons] = 0.0
    sol, out = h.solve(rhs)
    sb = osolver.BlockGMGStokes(ob)
    sb.A_block_is_symmetric = False
    want = sb.solve(rhs, tol_rel=1e-7, max_its=0, n_expensive=100)
    assert out["converged"] and want["converged"]
    assert out["iterations_expensive"] == want["iterations_expensive"]
    assert abs(out["n_A_expensive"] - want["n_A_cg"]) <= 2
    assert rel(sol[:act.n_u], want["x"][:act.n_u]) < 1e-6
    h.close()
