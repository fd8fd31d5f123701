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


def test_free_surface_face_term():
    """The boundary-face stabilisation term (matrix_free_operators.cc:563-602), oracle side: for g parallel to the
    normal it is the symmetric negative semi-definite boundary mass term -|g| int (u.n)(v.n); for a constant u it integrates
    -(g.u) n over the face; only face DoFs are touched; the matrix form agrees with the matrix-free form."""
    from oracle import mesh, newton
    for dim in (2, 3):
        L = mesh.Level(dim, 2 if dim == 2 else 1)
        face_no = 2 * dim - 1                                     # the top face
        cells = newton.boundary_faces(L, face_no)
        assert len(cells) == L.n ** (dim - 1)
        nfq = 3 ** (dim - 1)
        rng = np.random.default_rng(dim)
        # (a) g = gamma n
        gamma = 3.7
        tab = np.zeros((len(cells), nfq, dim)); tab[..., dim - 1] = gamma
        none = np.zeros(0, dtype=np.int64)
        K = newton.free_surface_matrix(L, cells, face_no, tab, none)
        assert abs(K - K.T).max() < 1e-14
        v = rng.uniform(-1, 1, L.n_u)
        assert v @ (K @ v) <= 1e-14
        u = np.zeros(L.n_u); u[dim - 1::dim] = 1.0               # u = n everywhere: -gamma * (face area = 1) in total
        assert abs((K @ u).sum() + gamma) < 1e-13
        top = np.zeros(L.n_u, bool)
        top[np.unique(L.cell_dofs_u[cells])] = True
        coords = L.dof_coords_u()
        assert not np.any((K @ v)[coords[:, dim - 1] < 1.0 - 1e-12])          # only DoFs on the top boundary receive anything
        # (b) general g: matrix-free form = matrix form, constrained rows / columns are skipped
        tab = rng.uniform(-1, 1, (len(cells), nfq, dim))
        cons = L.constrained_u("tangential")
        K = newton.free_surface_matrix(L, cells, face_no, tab, cons)
        y = newton.free_surface_apply_add(L, cells, face_no, tab, cons, v, np.zeros(L.n_u))
        assert np.abs(K @ v - y).max() < 1e-14
        assert not np.any(y[cons]) and abs(K[:, cons]).max() == 0.0


@pytest.mark.parametrize("dim,R", [(2, 2), (3, 1)])
@pytest.mark.parametrize("deviatoric", [False, True])
def test_newton_operator_is_the_jacobian_of_the_picard_operator(dim, R, deviatoric):
    """An anchor that does not depend on the restated formulas: with the tables filled as the reference fills them
    (stokes_matrix_free_local_smoothing.cc:436-445: strain_rate_table = eps(u0), newton_factor_wrt_strain_rate_table =
    scaling * d eta / d eps at u0) the unsymmetrised Newton operator must be the derivative of the Picard operator
    F(u) = A(eta(eps(u))) u at u0 -- checked by central differences for a power-law viscosity
    eta = eta0 (e_II^2 + e_min^2)^((1/n - 1)/2), e_II^2 = eps':eps'/2 (deviatoric) or eps:eps/2; and the symmetrised operator must be
    the symmetric part (J + J^T)/2 of that derivative's Newton term added to the Picard matrix."""
    lev = mesh.Level(dim, R)
    rng = np.random.default_rng(11)
    none = np.zeros(0, dtype=np.int64)
    nq, ns = 3 ** dim, dim * (dim + 1) // 2
    t = newton.fe.tensor_tables(dim)
    G = t["dphi2"] / lev.h[None, None, :]
    eta0, n_exp, e_min = 1.3, 3.0, 0.05
    I = np.eye(dim)

    def strain(xu):
        u = xu[lev.cell_dofs_u].reshape(lev.n_cells, dim, nq)
        g = np.einsum("kci,qid->kqcd", u, G)
        return 0.5 * (g + np.swapaxes(g, -1, -2))

    def viscosity(eps):
        e = eps - (np.trace(eps, axis1=-2, axis2=-1) / 3.0)[..., None, None] * I if deviatoric else eps
        s = 0.5 * np.einsum("kqcd,kqcd->kq", e, e) + e_min ** 2
        eta = eta0 * s ** (0.5 * (1.0 / n_exp - 1.0))
        # d eta / d eps_ij = eta0 (1/n - 1)/2 s^((1/n - 1)/2 - 1) * d s / d eps_ij;  d s / d eps = e : d e / d eps = e - tr(e)/3 I
        # for e = eps - tr(eps)/3 I (tr(e) = 0 in 3D only: the "deviator" keeps the reference's 1/3 in 2D as well), = eps otherwise
        ds = e - (np.trace(e, axis1=-2, axis2=-1) / 3.0)[..., None, None] * I if deviatoric else e
        deta = (eta0 * 0.5 * (1.0 / n_exp - 1.0) * s ** (0.5 * (1.0 / n_exp - 1.0) - 1.0))[..., None, None] * ds
        return eta, deta

    def picard(xu):                                    # F(u) = A(eta(eps(u))) u, velocity rows, no constraints
        eta, _ = viscosity(strain(xu))
        yu, _ = newton.stokes_apply(lev, eta, 1.0, none, xu, np.zeros(lev.n_p), deviatoric)
        return yu

    def unroll(tensor):                                # (cells, q, dim, dim) -> (cells, q, n_sym) in the tables' order
        return np.stack([tensor[..., i, j] for (i, j) in newton.SYM_INDEX[dim]], axis=-1)

    u0, du = rng.uniform(-1, 1, lev.n_u), rng.uniform(-1, 1, lev.n_u)
    eps0 = strain(u0)
    eta_q, deta = viscosity(eps0)
    nd = newton.NewtonData(dim, enable_newton_derivatives=True, symmetrize_newton_system=False, average_newton_factors=False,
                           strain_rate_table=unroll(eps0), newton_factor_wrt_strain_rate_table=unroll(deta),
                           newton_factor_wrt_pressure_table=np.zeros((lev.n_cells, nq)))
    j_du, _ = newton.stokes_apply(lev, eta_q, 1.0, none, du, np.zeros(lev.n_p), deviatoric, nd)
    h = 1e-5
    fd = (picard(u0 + h * du) - picard(u0 - h * du)) / (2.0 * h)
    assert np.linalg.norm(j_du - fd) <= 1e-7 * np.linalg.norm(fd), np.linalg.norm(j_du - fd) / np.linalg.norm(fd)
    # the derivative term matters (the Picard matrix alone is not the Jacobian)
    a_du, _ = newton.stokes_apply(lev, eta_q, 1.0, none, du, np.zeros(lev.n_p), deviatoric)
    assert np.linalg.norm(a_du - fd) > 1e-2 * np.linalg.norm(fd)
    # symmetrised system = Picard matrix + symmetric part of the Newton term: x.(J_s y) = x.(A y) + (x.(N y) + y.(N x)) / 2
    x, y = rng.uniform(-1, 1, lev.n_u), rng.uniform(-1, 1, lev.n_u)
    op = lambda v, data: newton.stokes_apply(lev, eta_q, 1.0, none, v, np.zeros(lev.n_p), deviatoric, data)[0]
    n_y, n_x = op(y, nd) - op(y, None), op(x, nd) - op(x, None)
    nd.symmetrize_newton_system = True
    lhs = x @ op(y, nd)
    rhs = x @ op(y, None) + 0.5 * (x @ n_y + y @ n_x)
    assert abs(lhs - rhs) <= 1e-12 * abs(rhs)
