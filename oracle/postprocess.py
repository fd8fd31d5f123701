"""Post-solve pieces used to pin the oracle against the reference's goldens (TEST INFRASTRUCTURE).

* pressure normalisation "surface" (ASPECT default): shift p so its mean over the top
  boundary equals the surface pressure 0 (/root/reference/source/simulator/helper_functions.cc,
  normalize_pressure).
* SolCx error norms: /root/reference/benchmarks/solcx/solcx.h:3129-3168
  (VectorTools::integrate_difference with QGauss(degree+2); the L1 norm of a vector field
  integrates sum_c |e_c|), analytic solution = the reference's own `_Velic_solCx`
  compiled into oracle/_ref/libsolcx.so by oracle/Makefile.
"""
import ctypes
import os
import numpy as np
from . import fe

_HERE = os.path.dirname(os.path.abspath(__file__))


def solcx_exact(points, eta_b=1e6):
    lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libsolcx.so"))
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((len(pts), 3))
    lib.solcx_exact(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(eta_b), ctypes.c_int(len(pts)),
                    out.ctypes.data_as(ctypes.c_void_p))
    return out


def normalize_pressure_surface(level, p):
    """Subtract the mean of the Q1 pressure over the top boundary (last coordinate = max)."""
    dim = level.dim
    g = level.node_grid_p
    top = np.where(g[:, dim - 1] == level.n)[0]
    # exact integral of the piecewise (bi)linear trace: tensor trapezoid weights
    wgt = np.ones(len(top))
    for d in range(dim - 1):
        edge = (g[top, d] == 0) | (g[top, d] == level.n)
        wgt *= np.where(edge, 0.5, 1.0)
    mean = (wgt * p[top]).sum() / wgt.sum()
    return p - mean, mean


def solcx_errors(level, u, p, eta_b=1e6):
    """(u_L1, p_L1, u_L2, p_L2) as printed by SolCxPostprocessor."""
    dim = 2
    xg, wg = fe.gauss(4)
    pts1 = [(a, b) for b in range(4) for a in range(4)]
    xq = np.array([[xg[a], xg[b]] for a, b in pts1])
    wq = np.array([wg[a] * wg[b] for a, b in pts1])
    v2x, _ = fe.q2_shape(xq[:, 0]); v2y, _ = fe.q2_shape(xq[:, 1])
    v1x, _ = fe.q1_shape(xq[:, 0]); v1y, _ = fe.q1_shape(xq[:, 1])
    phi2 = np.stack([v2x[:, i] * v2y[:, j] for j in range(3) for i in range(3)], axis=1)   # (q, 9)
    phi1 = np.stack([v1x[:, i] * v1y[:, j] for j in range(2) for i in range(2)], axis=1)
    X = (level.cells[:, None, :] + xq[None, :, :]) * level.h                               # (cells, q, 2)
    exact = solcx_exact(X.reshape(-1, 2), eta_b).reshape(level.n_cells, len(wq), 3)
    uh = np.stack([np.einsum("qi,ei->eq", phi2, u[level.cell_dofs_u[:, c * 9:(c + 1) * 9]]) for c in range(2)], axis=2)
    ph = np.einsum("qi,ei->eq", phi1, p[level.cell_dofs_p])
    JxW = wq[None, :] * level.cell_volume()
    eu = uh - exact[:, :, :2]
    ep = ph - exact[:, :, 2]
    u_l1 = (np.abs(eu).sum(axis=2) * JxW).sum()
    p_l1 = (np.abs(ep) * JxW).sum()
    u_l2 = np.sqrt(((eu ** 2).sum(axis=2) * JxW).sum())
    p_l2 = np.sqrt(((ep ** 2) * JxW).sum())
    return u_l1, p_l1, u_l2, p_l2


def normalize_pressure_volume(level, p):
    """`Pressure normalization = volume` (benchmarks/nsinker/nsinker_gmg.prm): subtract the volume mean of the
    Q1 pressure (exact integral: cell volume times the mean of the cell's nodal values)."""
    mean = p[level.cell_dofs_p].mean(axis=1).sum() * level.cell_volume() / float(np.prod(level.extents))
    return p - mean, mean
