"""CPU checks of the oracle's complete Stokes cell operation (Newton / dilation / compressible branches,
matrix_free_operators.cc:376-547) against anchors that do not depend on it."""
import numpy as np
import pytest

from oracle import assembled, mesh, newton


def random_newton_data(level, rng, **flags):
    dim, nc, nq = level.dim, level.n_cells, 3 ** level.dim
    ns = dim * (dim + 1) // 2
    return newton.NewtonData(
        dim, strain_rate_table=rng.uniform(-1, 1, (nc, nq, ns)), newton_factor_wrt_strain_rate_table=rng.uniform(-1, 1, (nc, nq, ns)),
        newton_factor_wrt_pressure_table=rng.uniform(-1, 1, (nc, nq)), dilation_lhs_term_table=rng.uniform(0, 1, (nc, nq)),
        dilation_derivative_wrt_strain_rate_table=rng.uniform(-1, 1, (nc, nq, ns)),
        dilation_derivative_wrt_pressure_table=rng.uniform(-1, 1, (nc, nq)), **flags)


@pytest.mark.parametrize("dim,R", [(2, 2), (3, 1)])
@pytest.mark.parametrize("compressible", [False, True])
@pytest.mark.parametrize("perq", [False, True])
def test_flags_off_equals_assembled(dim, R, compressible, perq):
    lev = mesh.Level(dim, R)
    rng = np.random.default_rng(3)
    eta = np.exp(rng.uniform(-2, 2, (lev.n_cells, 3 ** dim if perq else 1)))
    cons = lev.constrained_u("zero")
    ps = 1.7
    K, A, BT = assembled.stokes_matrix(lev, eta, ps, cons, compressible)
    x = rng.uniform(-1, 1, lev.n_u + lev.n_p)
    yu, yp = newton.stokes_apply(lev, eta, ps, cons, x[:lev.n_u], x[lev.n_u:], compressible)
    want = K @ x
    assert np.linalg.norm(np.concatenate([yu, yp]) - want) <= 1e-13 * np.linalg.norm(want)
    ya = newton.A_apply(lev, eta, cons, x[:lev.n_u], compressible)
    assert np.linalg.norm(ya - A @ x[:lev.n_u]) <= 1e-13 * np.linalg.norm(ya)


@pytest.mark.parametrize("dim,R", [(2, 2), (3, 1)])
@pytest.mark.parametrize("average", [False, True])
@pytest.mark.parametrize("deviatoric", [False, True])
def test_symmetrized_newton_system_is_symmetric(dim, R, average, deviatoric):
    lev = mesh.Level(dim, R)
    rng = np.random.default_rng(5)
    eta = np.exp(rng.uniform(-1, 1, (lev.n_cells, 3 ** dim)))
    none = np.zeros(0, dtype=np.int64)

    def op(nd, x):
        yu, yp = newton.stokes_apply(lev, eta, 1.0, none, x, np.zeros(lev.n_p), deviatoric, nd)
        return yu

    x, y = rng.uniform(-1, 1, lev.n_u), rng.uniform(-1, 1, lev.n_u)
    nd = random_newton_data(lev, np.random.default_rng(6), enable_newton_derivatives=True, symmetrize_newton_system=True,
                            average_newton_factors=average)
    a, b = y @ op(nd, x), x @ op(nd, y)
    assert abs(a - b) <= 1e-12 * max(abs(a), abs(b))
    nd.symmetrize_newton_system = False
    a, b = y @ op(nd, x), x @ op(nd, y)
    assert abs(a - b) > 1e-6 * max(abs(a), abs(b))          # (the unsymmetrised Newton matrix is not symmetric)


def test_dilation_changes_only_what_it_should():
    """enable_prescribed_dilation without Newton derivatives: deviatoric stress + the -ps^2 dilation_lhs p term in the
    pressure rows (matrix_free_operators.cc:464-479); the A block becomes deviatoric (:999-1005)."""
    lev = mesh.Level(2, 2)
    rng = np.random.default_rng(8)
    eta = np.exp(rng.uniform(-1, 1, (lev.n_cells, 1)))
    cons = lev.constrained_u("zero")
    ps = 2.0
    nd = random_newton_data(lev, rng, enable_prescribed_dilation=True)
    x = rng.uniform(-1, 1, lev.n_u + lev.n_p)
    yu, yp = newton.stokes_apply(lev, eta, ps, cons, x[:lev.n_u], x[lev.n_u:], False, nd)
    cu, cp = newton.stokes_apply(lev, eta, ps, cons, x[:lev.n_u], x[lev.n_u:], True)
    assert np.allclose(yu, cu, rtol=0, atol=1e-13 * np.abs(cu).max())
    # pressure rows differ by the (ps^2 dilation_lhs)-weighted Q1 mass matrix applied to p
    M = assembled.assemble_mass(lev, 1.0 / nd.dilation_lhs_term_table, ps) if hasattr(assembled, "assemble_mass") else None
    if M is not None:
        assert np.allclose(cp - yp, M @ x[lev.n_u:], rtol=0, atol=1e-12 * np.abs(yp).max())
    assert np.allclose(newton.A_apply(lev, eta, cons, x[:lev.n_u], False, True), newton.A_apply(lev, eta, cons, x[:lev.n_u], True))
