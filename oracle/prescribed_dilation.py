"""The reference's `tests/prescribed_dilation[_gmg]` benchmark (oracle; test infrastructure).

/root/reference/tests/prescribed_dilation.cc: box [0,2]^2, constant viscosity eta = 10, `Enable prescribed dilation` with
dilation_rhs_term = x, dilation_lhs_term = 0 (:127-131), `Enable additional Stokes RHS` with the force (-eta (1 - cos y),
eta sin x) (:121-126), analytic solution u = (cos y, sin x + x y), p = -2/3 eta (x - 1) (:36-50) prescribed on all four
boundaries, zero gravity, `Pressure normalization = volume`, errors with QGauss(degree + 2) (:218).
Right-hand side: source/simulator/assemblers/stokes.cc:387-412 (force, and -pressure_scaling * dilation_rhs_term * phi_p).
Operators: the ACTIVE Stokes operator and A block are deviatoric (matrix_free_operators.cc:442, :1000); the LEVEL operators are
not -- stokes_matrix_free_local_smoothing.cc:312 sets enable_prescribed_dilation = false on every level and the material is
incompressible -- which is what the golden's 35 / 36 iterations pin (26 / 27 with deviatoric level operators).
"""
import numpy as np

from . import assembled, fe

ETA = 10.0
EXTENT = 2.0


def velocity(x):
    return np.stack([np.cos(x[..., 1]), np.sin(x[..., 0]) + x[..., 0] * x[..., 1]], axis=-1)


def pressure(x):
    return -2.0 / 3.0 * ETA * (x[..., 0] - 1.0)


def problem(level):
    """(viscosity_q, pressure_scaling, rhs [u | p] before the boundary correction, u0 = boundary values on the constrained DoFs)."""
    xq = level.quadrature_points()
    eta_q = np.full(xq.shape[:2], ETA)
    ps = ETA / (0.01 * EXTENT)                                   # helper_functions.cc:253-348, box length scale 0.01 * X extent
    x, y = xq[..., 0], xq[..., 1]
    force = np.stack([-ETA * (1.0 - np.cos(y)), ETA * np.sin(x)], axis=-1)
    f_u = assembled.rhs_force(level, force)
    t = fe.tensor_tables(2)
    f_p = np.zeros(level.n_p)
    np.add.at(f_p, np.asarray(level.cell_dofs_p, dtype=np.int64),
              np.einsum("eq,qi->ei", -ps * x * (t["w"] * float(np.prod(level.h)))[None, :], t["phi1"]))
    cons = np.asarray(level.constrained_u("zero"), dtype=np.int64)
    u0 = np.zeros(level.n_u)
    v = velocity(level.dof_coords_u())                           # (n_u, 2): every DoF carries its node's coordinates
    u0[cons] = v[cons, cons % 2]
    return eta_q, ps, np.concatenate([f_u, f_p]), u0


def errors(level, u, p):
    """(u_L2, p_L2) as printed by MyPostprocessor (QGauss(4) per cell)."""
    xg, wg = fe.gauss(4)
    pts = [(a, b) for b in range(4) for a in range(4)]
    xr = np.array([[xg[a], xg[b]] for a, b in pts])
    wq = np.array([wg[a] * wg[b] for a, b in pts])
    v2x, _ = fe.q2_shape(xr[:, 0]); v2y, _ = fe.q2_shape(xr[:, 1])
    v1x, _ = fe.q1_shape(xr[:, 0]); v1y, _ = fe.q1_shape(xr[:, 1])
    phi2 = np.stack([v2x[:, i] * v2y[:, j] for j in range(3) for i in range(3)], axis=1)
    phi1 = np.stack([v1x[:, i] * v1y[:, j] for j in range(2) for i in range(2)], axis=1)
    X = (np.asarray(level.cells)[:, None, :] + xr[None]) * level.h
    cu, cp = np.asarray(level.cell_dofs_u, dtype=np.int64), np.asarray(level.cell_dofs_p, dtype=np.int64)
    uh = np.stack([np.einsum("qi,ei->eq", phi2, u[cu[:, c * 9:(c + 1) * 9]]) for c in range(2)], axis=2)
    ph = np.einsum("qi,ei->eq", phi1, p[cp])
    JxW = wq[None] * float(np.prod(level.h))
    return (float(np.sqrt((((uh - velocity(X)) ** 2).sum(2) * JxW).sum())),
            float(np.sqrt((((ph - pressure(X)) ** 2) * JxW).sum())))
