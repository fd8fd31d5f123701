"""Independently assembled sparse Q2-Q1 operators (oracle; test infrastructure).

Built from the textbook weak form the matrix-based assembler uses
(/root/reference/source/simulator/assemblers/stokes.cc:417-433):
    A   = int 2 eta eps(phi_i):eps(phi_j)           [- (2 eta/3) div div if compressible]
    B^T = - ps int div(phi_i) psi_j
    M   = int (ps^2/eta) psi_i psi_j
with the 3^dim Gauss rule.  Constrained rows/columns are treated the way
MatrixFreeOperators::Base does: reads of constrained DoFs give 0,
writes are skipped, and the constrained diagonal is 1.
Used to (1) validate the matrix-free restatement and the CUDA kernels on small meshes,
(2) reproduce the reference's golden iteration counts without any matrix-free code.
"""
import numpy as np
import scipy.sparse as sp
from . import fe


def _eta_q(eta, n_cells, nq):
    eta = np.asarray(eta, float)
    if eta.ndim == 1:
        eta = eta[:, None]
    if eta.shape[1] == 1:
        eta = np.repeat(eta, nq, axis=1)
    assert eta.shape == (n_cells, nq)
    return eta


def local_gradients(level):
    """Physical shape gradients of the vector-valued Q2 basis, one Cartesian cell.
    returns G[q, I, a, b] = d_b (phi_I)_a with I = c*3^dim + node (component-major)."""
    dim = level.dim
    t = fe.tensor_tables(dim)
    nn = 3 ** dim
    G = np.zeros((len(t["w"]), dim * nn, dim, dim))
    for c in range(dim):
        G[:, c * nn:(c + 1) * nn, c, :] = t["dphi2"] / level.h[None, None, :]
    return G, t


def _curved(level):
    return getattr(level, "jinv", None) is not None


def _phys_gradients(level, t):
    """Curved cells: g[e, q, i, d] = sum_r dphi2[q, i, r] jinv[e, q, r, d] (scalar Q2 shape gradients in real space)."""
    return np.einsum("qir,eqrd->eqid", t["dphi2"], level.jinv)


def _assemble_A_curved(level, eta, constrained, compressible):
    """2 eta eps(phi_I) : eps(phi_J) with I = (c, i), J = (c', j):
    block (c, c') = sum_q eta JxW [ delta_cc' grad phi_i . grad phi_j + d_c' phi_i d_c phi_j - (2/3) d_c phi_i d_c' phi_j ]."""
    dim = level.dim
    t = fe.tensor_tables(dim)
    nn = 3 ** dim
    eta = _eta_q(eta, level.n_cells, len(t["w"]))
    g = _phys_gradients(level, t)
    coef = eta * level.jxw
    Ke = np.zeros((level.n_cells, dim * nn, dim * nn))
    lap = np.einsum("eq,eqid,eqjd->eij", coef, g, g)
    for c in range(dim):
        for c2 in range(dim):
            blk = np.einsum("eq,eqi,eqj->eij", coef, g[..., c2], g[..., c])
            if compressible:
                blk = blk - (2.0 / 3.0) * np.einsum("eq,eqi,eqj->eij", coef, g[..., c], g[..., c2])
            if c == c2:
                blk = blk + lap
            Ke[:, c * nn:(c + 1) * nn, c2 * nn:(c2 + 1) * nn] = blk
    return _scatter(Ke, level.cell_dofs_u, level.cell_dofs_u, level.n_u, level.n_u, constrained, constrained, True)


def assemble_A(level, eta, constrained=None, compressible=False):
    if _curved(level):
        return _assemble_A_curved(level, eta, constrained, compressible)
    dim = level.dim
    G, t = local_gradients(level)
    nq = len(t["w"])
    eta = _eta_q(eta, level.n_cells, nq)
    E = 0.5 * (G + np.swapaxes(G, 2, 3))                       # sym grad
    if compressible:
        tr = np.trace(E, axis1=2, axis2=3)
        Edev = E - tr[..., None, None] * np.eye(dim) / 3.0
    else:
        Edev = E
    # K_q[I,J] = 2 * E_I : Edev_J
    Kq = 2.0 * np.einsum("qiab,qjab->qij", E, Edev)
    coef = eta * (t["w"] * level.cell_volume())[None, :]         # (cells, q)
    Ke = np.einsum("eq,qij->eij", coef, Kq)
    return _scatter(Ke, level.cell_dofs_u, level.cell_dofs_u, level.n_u, level.n_u, constrained, constrained, True)


def assemble_BT(level, pressure_scaling=1.0, constrained_u=None):
    """(n_u x n_p): u_dst = -ps int p div v."""
    dim = level.dim
    if _curved(level):
        t = fe.tensor_tables(dim)
        g = _phys_gradients(level, t)
        Ke = np.concatenate([-pressure_scaling * np.einsum("eq,eqi,qj->eij", level.jxw, g[..., c], t["phi1"]) for c in range(dim)], axis=1)
        return _scatter(Ke, level.cell_dofs_u, level.cell_dofs_p, level.n_u, level.n_p, constrained_u, None, False)
    G, t = local_gradients(level)
    div = np.trace(G, axis1=2, axis2=3)                           # (q, I)
    coef = t["w"] * level.cell_volume()
    Ke1 = -pressure_scaling * np.einsum("q,qi,qj->ij", coef, div, t["phi1"])
    Ke = np.broadcast_to(Ke1, (level.n_cells,) + Ke1.shape)
    return _scatter(Ke, level.cell_dofs_u, level.cell_dofs_p, level.n_u, level.n_p, constrained_u, None, False)


def assemble_mass(level, eta, pressure_scaling=1.0):
    dim = level.dim
    t = fe.tensor_tables(dim)
    nq = len(t["w"])
    eta = _eta_q(eta, level.n_cells, nq)
    jxw = level.jxw if _curved(level) else (t["w"] * level.cell_volume())[None, :]
    coef = (pressure_scaling ** 2 / eta) * jxw
    Ke = np.einsum("eq,qi,qj->eij", coef, t["phi1"], t["phi1"])
    return _scatter(Ke, level.cell_dofs_p, level.cell_dofs_p, level.n_p, level.n_p, None, None, False)


def _scatter(Ke, rows, cols, nr, nc, crow, ccol, unit_diag):
    ne, lr, lc = Ke.shape
    R = np.repeat(rows[:, :, None], lc, axis=2).ravel()
    C = np.repeat(cols[:, None, :], lr, axis=1).ravel()
    V = np.ascontiguousarray(Ke).ravel().copy()
    if crow is not None and len(crow):
        mask = np.zeros(nr, bool)
        mask[crow] = True
        V[mask[R]] = 0.0
    if ccol is not None and len(ccol):
        mask = np.zeros(nc, bool)
        mask[ccol] = True
        V[mask[C]] = 0.0
    M = sp.coo_matrix((V, (R, C)), shape=(nr, nc)).tocsr()
    if unit_diag and crow is not None and len(crow):
        M = M + sp.coo_matrix((np.ones(len(crow)), (crow, crow)), shape=(nr, nc)).tocsr()
    M.eliminate_zeros()
    return M


def stokes_matrix(level, eta, pressure_scaling=1.0, constrained_u=None, compressible=False):
    """[A B^T; B 0] with the MatrixFreeOperators::Base constraint convention."""
    A = assemble_A(level, eta, constrained_u, compressible)
    BT = assemble_BT(level, pressure_scaling, constrained_u)
    return sp.bmat([[A, BT], [BT.T, None]], format="csr"), A, BT


def rhs_gravity(level, rho_q, gravity):
    """f_i = int rho g . phi_i (assemblers/stokes.cc:387-396), rho_q (cells, nq); gravity (dim,)."""
    dim = level.dim
    t = fe.tensor_tables(dim)
    nn = 3 ** dim
    coef = rho_q * (level.jxw if _curved(level) else (t["w"] * level.cell_volume())[None, :])
    loc = np.einsum("eq,qi->ei", coef, t["phi2"])                # (cells, nodes)
    f = np.zeros(level.n_u)
    for c in range(dim):
        np.add.at(f, level.cell_dofs_u[:, c * nn:(c + 1) * nn], gravity[c] * loc)
    return f


def rhs_force(level, force_q):
    """f_i = int force . phi_i for a force density given at the quadrature points, force_q (cells, nq, dim)
    (e.g. rho * g with a radial gravity vector; assemblers/stokes.cc:387-396)."""
    dim = level.dim
    t = fe.tensor_tables(dim)
    nn = 3 ** dim
    jxw = level.jxw if _curved(level) else np.broadcast_to((t["w"] * level.cell_volume())[None, :], force_q.shape[:2])
    f = np.zeros(level.n_u)
    for c in range(dim):
        loc = np.einsum("eq,qi->ei", force_q[..., c] * jxw, t["phi2"])
        np.add.at(f, level.cell_dofs_u[:, c * nn:(c + 1) * nn], loc)
    return f
