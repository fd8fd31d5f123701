"""Oracle (test infrastructure): the COMPLETE cell operation of the reference's matrix-free Stokes operator, including
the Newton-derivative, prescribed-dilation and compressible branches, restated with numpy over all cells at once.

Follows /root/reference/source/simulator/solver/matrix_free_operators.cc
    StokesOperator::local_apply         :376-547  (every branch, same variable names)
    ABlockOperator::inner_cell_operation :974-1009 (deviatoric stress if compressible or prescribed dilation)
and the table layout of OperatorCellData (include/aspect/simulator/solver/matrix_free_operators.h:91-208):
SymmetricTensor<2,dim> tables are stored with deal.II's unrolled component order 00, 11, (22,) 01, (02, 12).

Constraint handling is that of MatrixFreeOperators::Base: constrained source entries are read as 0
(unless `plain`), nothing is added to constrained rows, and dst[c] = src[c] on them.

Independent anchors (tests/test_newton_oracle.py): with every extra flag off this reproduces the assembled sparse
operator of oracle/assembled.py to 1e-13 (incompressible and compressible); with `symmetrize_newton_system` (and no
pressure derivative) the operator is symmetric in all four averaged / deviatoric combinations, without it it is not.
"""
import numpy as np

from . import fe

SYM_INDEX = {2: [(0, 0), (1, 1), (0, 1)], 3: [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]}


def sym_to_full(t, dim):
    """(..., n_sym) unrolled symmetric tensors -> (..., dim, dim)."""
    out = np.zeros(t.shape[:-1] + (dim, dim))
    for k, (i, j) in enumerate(SYM_INDEX[dim]):
        out[..., i, j] = t[..., k]
        out[..., j, i] = t[..., k]
    return out


class NewtonData:
    """OperatorCellData fields beyond the viscosity (matrix_free_operators.h:128-196); tables are [cell][q][...]."""

    def __init__(self, dim, enable_newton_derivatives=False, symmetrize_newton_system=False, average_newton_factors=False,
                 enable_prescribed_dilation=False, strain_rate_table=None, newton_factor_wrt_strain_rate_table=None,
                 newton_factor_wrt_pressure_table=None, dilation_lhs_term_table=None,
                 dilation_derivative_wrt_strain_rate_table=None, dilation_derivative_wrt_pressure_table=None):
        self.dim = dim
        self.enable_newton_derivatives = enable_newton_derivatives
        self.symmetrize_newton_system = symmetrize_newton_system
        self.average_newton_factors = average_newton_factors
        self.enable_prescribed_dilation = enable_prescribed_dilation
        self.strain_rate_table = strain_rate_table
        self.newton_factor_wrt_strain_rate_table = newton_factor_wrt_strain_rate_table
        self.newton_factor_wrt_pressure_table = newton_factor_wrt_pressure_table
        self.dilation_lhs_term_table = dilation_lhs_term_table
        self.dilation_derivative_wrt_strain_rate_table = dilation_derivative_wrt_strain_rate_table
        self.dilation_derivative_wrt_pressure_table = dilation_derivative_wrt_pressure_table


def _gather(level, constrained_u, xu, xp, plain):
    dim = level.dim
    nn = 3 ** dim
    xu_r = xu.copy()
    if not plain and len(constrained_u):
        xu_r[constrained_u] = 0.0
    u = xu_r[level.cell_dofs_u].reshape(level.n_cells, dim, nn)           # [cell][c][i]
    p = None if xp is None else xp[level.cell_dofs_p]                     # [cell][a]
    return u, p


def _tables(level):
    t = fe.tensor_tables(level.dim)
    G = t["dphi2"] / level.h[None, None, :]                                # physical gradients [q][i][d]
    JxW = t["w"] * level.cell_volume()                                     # [q]
    return t, G, JxW


def stokes_apply(level, eta, ps, constrained_u, xu, xp, is_compressible=False, nd=None, plain=False):
    """StokesOperator::vmult with every branch of local_apply.  eta: (cells, 1 or n_q).  Returns (yu, yp)."""
    dim = level.dim
    nd = nd or NewtonData(dim)
    t, G, JxW = _tables(level)
    u, p = _gather(level, constrained_u, xu, xp, plain)
    grad = np.einsum("kci,qid->kqcd", u, G)                                # d u_c / d x_d at q
    sym_grad_u = 0.5 * (grad + np.swapaxes(grad, -1, -2))
    div_u = np.trace(sym_grad_u, axis1=-2, axis2=-1)                       # [cell][q]
    val_p = np.einsum("ka,qa->kq", p, t["phi1"])
    nq = len(JxW)
    viscosity_x_2 = 2.0 * (eta if eta.shape[1] == nq else np.repeat(eta, nq, axis=1))
    I = np.eye(dim)
    deviatoric = is_compressible or nd.enable_prescribed_dilation

    pressure_terms = -ps * div_u
    if nd.enable_prescribed_dilation:
        pressure_terms = pressure_terms - ps * ps * nd.dilation_lhs_term_table * val_p
    velocity_terms = viscosity_x_2[..., None, None] * sym_grad_u
    velocity_terms = velocity_terms - (ps * val_p)[..., None, None] * I
    if deviatoric:
        velocity_terms = velocity_terms - (viscosity_x_2 / 3.0 * div_u)[..., None, None] * I

    if nd.enable_newton_derivatives:
        eps = sym_to_full(nd.strain_rate_table, dim)                        # strain_rate_table
        deta = sym_to_full(nd.newton_factor_wrt_strain_rate_table, dim)     # newton_factor_wrt_strain_rate_table
        detadp = nd.newton_factor_wrt_pressure_table
        tr_eps = np.trace(eps, axis1=-2, axis2=-1)
        if nd.average_newton_factors:
            deta_deps_times_sym_grad_u = np.einsum("kqcd,kqcd->k", deta, sym_grad_u)[:, None] * np.ones(nq)
            deta_dp_times_p = (detadp * val_p).sum(axis=1)[:, None] * np.ones(nq)
            eps_times_sym_grad_u = theta_times_div_u = None
            if nd.symmetrize_newton_system:
                eps_JxW = np.einsum("kqcd,kqcd,q->k", eps, sym_grad_u, JxW)
                eps_times_sym_grad_u = eps_JxW[:, None] / JxW[None, :]
                if deviatoric:
                    theta_JxW = np.einsum("kq,kq,q->k", tr_eps, div_u, JxW)
                    theta_times_div_u = theta_JxW[:, None] / JxW[None, :]
        else:
            deta_deps_times_sym_grad_u = np.einsum("kqcd,kqcd->kq", deta, sym_grad_u)
            deta_dp_times_p = detadp * val_p
            eps_times_sym_grad_u = np.einsum("kqcd,kqcd->kq", eps, sym_grad_u) if nd.symmetrize_newton_system else None
            theta_times_div_u = tr_eps * div_u
        if nd.symmetrize_newton_system:
            velocity_terms = velocity_terms + eps * deta_deps_times_sym_grad_u[..., None, None] + deta * eps_times_sym_grad_u[..., None, None]
        else:
            velocity_terms = velocity_terms + 2.0 * eps * deta_deps_times_sym_grad_u[..., None, None]
        velocity_terms = velocity_terms + 2.0 * eps * deta_dp_times_p[..., None, None]
        if deviatoric:
            if nd.symmetrize_newton_system:
                velocity_terms = velocity_terms - (1.0 / 3.0) * deta * theta_times_div_u[..., None, None]
                velocity_terms = velocity_terms - ((1.0 / 3.0) * tr_eps * deta_deps_times_sym_grad_u)[..., None, None] * I
            else:
                velocity_terms = velocity_terms - ((2.0 / 3.0) * tr_eps * deta_deps_times_sym_grad_u)[..., None, None] * I
            if nd.enable_prescribed_dilation:
                ddil = sym_to_full(nd.dilation_derivative_wrt_strain_rate_table, dim)
                pressure_terms = pressure_terms + (np.einsum("kqcd,kqcd->kq", ddil, sym_grad_u)
                                                   + nd.dilation_derivative_wrt_pressure_table * ps * val_p) * ps

    out_u = np.einsum("kqcd,qid,q->kci", velocity_terms, G, JxW)            # submit_symmetric_gradient + integrate
    out_p = np.einsum("kq,qa,q->ka", pressure_terms, t["phi1"], JxW)
    yu = np.zeros(level.n_u)
    np.add.at(yu, level.cell_dofs_u.ravel(), out_u.reshape(level.n_cells, -1).ravel())
    yp = np.zeros(level.n_p)
    np.add.at(yp, level.cell_dofs_p.ravel(), out_p.ravel())
    if len(constrained_u):
        yu[constrained_u] = 0.0 if plain else xu[constrained_u]
    return yu, yp


def A_apply(level, eta, constrained_u, xu, is_compressible=False, enable_prescribed_dilation=False):
    """ABlockOperator::vmult (inner_cell_operation :974-1009)."""
    dim = level.dim
    t, G, JxW = _tables(level)
    u, _ = _gather(level, constrained_u, xu, None, False)
    grad = np.einsum("kci,qid->kqcd", u, G)
    nq = len(JxW)
    viscosity_x_2 = 2.0 * (eta if eta.shape[1] == nq else np.repeat(eta, nq, axis=1))
    s = viscosity_x_2[..., None, None] * 0.5 * (grad + np.swapaxes(grad, -1, -2))
    if is_compressible or enable_prescribed_dilation:
        div = np.trace(s, axis1=-2, axis2=-1)
        s = s - (div / 3.0)[..., None, None] * np.eye(dim)
    out_u = np.einsum("kqcd,qid,q->kci", s, G, JxW)
    yu = np.zeros(level.n_u)
    np.add.at(yu, level.cell_dofs_u.ravel(), out_u.reshape(level.n_cells, -1).ravel())
    if len(constrained_u):
        yu[constrained_u] = xu[constrained_u]
    return yu


# ---- free-surface stabilisation faces (matrix_free_operators.cc:563-602) -------------------------------------------
def boundary_faces(level, face_no):
    """Cells of a box Level that have local face `face_no` (deal.II numbering x0, x1, y0, y1[, z0, z1]) on the boundary."""
    d, side = face_no // 2, face_no % 2
    want = 0 if side == 0 else level.n - 1
    return np.flatnonzero(level.cells[:, d] == want)


def free_surface_apply_add(level, cells, face_no, stabilization_term_table, constrained_u, xu, yu, plain=False):
    """StokesOperator::local_apply_boundary_face: on every free-surface boundary face
        yu_I -= int_face (g . u)(n . phi_I) dA
    with g = stabilization_term_table[face][q] (= density * time step * theta * |gravity| * g_hat, filled by
    stokes_matrix_free_local_smoothing.cc:555-612), n the outward normal, QGauss<dim-1>(degree + 1) = 3^(dim-1) face points
    (q lexicographic over the two face-tangential directions, lower direction fastest).  Source values are read like in the
    cell loop (constrained DoFs give 0 unless `plain`), nothing is added to constrained rows.
    cells: the cells whose local face `face_no` is a free-surface face; stabilization_term_table: (len(cells), 3^(dim-1), dim)."""
    dim = level.dim
    d, side = face_no // 2, face_no % 2
    tang = [e for e in range(dim) if e != d]
    normal = np.zeros(dim); normal[d] = -1.0 if side == 0 else 1.0
    area = float(np.prod([level.h[e] for e in tang]))
    # face nodes of the cell (lexicographic cell index) and the Q2 face shape functions at the face Gauss points
    lex = fe.lex_nodes(3, dim)
    face_nodes = [a for a, loc in enumerate(lex) if loc[d] == (0 if side == 0 else 2)]
    fq = fe.lex_nodes(3, dim - 1)
    phi = np.array([[np.prod([fe.N2[q[k], lex[a][tang[k]]] for k in range(dim - 1)]) for a in face_nodes] for q in fq])   # [q][face node]
    w = np.array([np.prod([fe.GAUSS_W[q[k]] for k in range(dim - 1)]) for q in fq]) * area
    nn = 3 ** dim
    xr = xu.copy()
    if not plain and len(constrained_u):
        xr[constrained_u] = 0.0
    mask = np.ones(level.n_u, bool)
    if len(constrained_u):
        mask[constrained_u] = False
    for k, cell in enumerate(cells):
        dofs = np.array([[level.cell_dofs_u[cell, c * nn + a] for a in face_nodes] for c in range(dim)])     # [c][face node]
        u_q = phi @ xr[dofs].T                                                  # [q][c]
        g_dot_u = (stabilization_term_table[k] * u_q).sum(axis=1)               # [q]
        val = -(g_dot_u * w)[:, None] * normal[None, :]                         # submit_value: -(g . u) n
        contrib = phi.T @ val                                                   # [face node][c]
        for c in range(dim):
            idx = dofs[c]
            np.add.at(yu, idx[mask[idx]], contrib[mask[idx], c])
    return yu


def free_surface_matrix(level, cells, face_no, stabilization_term_table, constrained_u):
    """The same term as a dense-assembled sparse matrix (rows: test functions), for checks."""
    import scipy.sparse as sp
    cols = []
    for j in range(level.n_u):
        e = np.zeros(level.n_u); e[j] = 1.0
        cols.append(free_surface_apply_add(level, cells, face_no, stabilization_term_table, constrained_u, e, np.zeros(level.n_u)))
    return sp.csr_matrix(np.array(cols).T)
