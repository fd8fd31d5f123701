"""Hollow-sphere benchmark (oracle; test infrastructure): analytic solution, material and error norms.

Restates /root/reference/benchmarks/hollow_sphere/hollow_sphere.cc for `Viscosity parameter` m:
  velocity :56-98, pressure :100-131, material (viscosity r^(m+1), density) :355-400,
  error norms (QGauss(degree + 2) = 4 points, L1 and L2 of velocity and pressure) :496-568,
with the spherical coordinates of source/utilities.cc:697-720 (r, phi = atan2(y, x), theta = acos(z / r)).
The reference's own test tests/hollow_sphere_gmg runs it on the 12-cell shell refined once (96 cells, 2,910 + 150 DoFs)
with `Stokes solver type = block GMG`: 33 iterations, errors 1.307365e-01, 1.699214e+00, 5.712572e-02, 1.076640e+00 --
the only reference golden with curved cells whose mesh this repository can rebuild up to the placement of the
mapping's interior points (see oracle/shell.py).
"""
import numpy as np

from . import fe

R1, R2, GAMMA, MU0, RHO0, GRAVITY = 0.5, 1.0, 1.0, 1.0, 1000.0, 1.0


def _alpha_beta(m):
    if m == -1:
        alpha = -GAMMA * (R2 ** 3 - R1 ** 3) / (R2 ** 3 * np.log(R1) - R1 ** 3 * np.log(R2))
        beta = -3 * GAMMA * (np.log(R2) - np.log(R1)) / (R1 ** 3 * np.log(R2) - R2 ** 3 * np.log(R1))
    else:
        alpha = GAMMA * (m + 1) * (R1 ** -3 - R2 ** -3) / (R1 ** (-m - 4) - R2 ** (-m - 4))
        beta = -3 * GAMMA * (R1 ** (m + 1) - R2 ** (m + 1)) / (R1 ** (m + 4) - R2 ** (m + 4))
    return alpha, beta


def _fg(r, m):
    alpha, beta = _alpha_beta(m)
    if m == -1:
        fr = alpha / (r * r) + beta * r
        gr = -2 / (r * r) * (alpha * np.log(r) + beta / 3 * r ** 3 + GAMMA)
    else:
        fr = alpha / r ** (m + 3) + beta * r
        gr = -2 / (r * r) * (-alpha / (m + 1) * r ** (-m - 1) + beta / 3 * r ** 3 + GAMMA)
    return fr, gr


def spherical(x):
    r = np.sqrt((x * x).sum(-1))
    phi = np.arctan2(x[..., 1], x[..., 0])
    phi = np.where(phi < 0.0, phi + 2.0 * np.pi, phi)
    theta = np.arccos(x[..., 2] / r)
    return r, phi, theta


def velocity(x, m=-1):
    r, phi, theta = spherical(x)
    fr, gr = _fg(r, m)
    v_r, v_theta, v_phi = gr * np.cos(theta), fr * np.sin(theta), fr * np.sin(theta)
    vx = np.sin(theta) * np.cos(phi) * v_r + np.cos(theta) * np.cos(phi) * v_theta - np.sin(phi) * v_phi
    vy = np.sin(theta) * np.sin(phi) * v_r + np.cos(theta) * np.sin(phi) * v_theta + np.cos(phi) * v_phi
    vz = np.cos(theta) * v_r - np.sin(theta) * v_theta
    return np.stack([vx, vy, vz], axis=-1)


def pressure(x, m=-1):
    r, phi, theta = spherical(x)
    fr, gr = _fg(r, m)
    mur = MU0 if m == -1 else MU0 * r ** (m + 1)
    hr = 2 / r * gr * mur if m == -1 else (m + 3) / r * gr * mur
    return hr * np.cos(theta) + RHO0 * GRAVITY * (R2 - r)


def material(x, m=-1):
    """(viscosity, density)."""
    r, phi, theta = spherical(x)
    alpha, beta = _alpha_beta(m)
    if m == -1:
        rho = -(alpha / r ** 4 * (8 * np.log(r) - 6) + 8.0 / 3.0 * beta / r + 8 * GAMMA / r ** 4) * np.cos(theta) + RHO0
    else:
        rho = -(2 * alpha * r ** -4 * (m + 3) / (m + 1) * (m - 1) - 2 * beta / 3 * (m - 1) * (m + 3) * r ** m
                - m * (m + 5) * 2 * GAMMA * r ** (m - 3)) * np.cos(theta) + RHO0
    return r ** (m + 1), rho


def _lagrange(nodes, x):
    n = len(nodes)
    V = np.ones((len(x), n)); D = np.zeros((len(x), n))
    for a in range(n):
        for b in range(n):
            if b != a:
                V[:, a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
        for b in range(n):
            if b != a:
                t = np.ones(len(x)) / (nodes[a] - nodes[b])
                for c in range(n):
                    if c not in (a, b):
                        t *= (x - nodes[c]) / (nodes[a] - nodes[c])
                D[:, a] += t
    return V, D


def _tensor3(A):
    """[q][k] tables of the tensor product basis from the 1D table A[q1][k1] (x fastest in q and k)."""
    n, k = A.shape
    return np.einsum("cl,bj,ai->cbalji", A, A, A).reshape(n ** 3, k ** 3)


def errors(level, sol_u, sol_p, m=-1):
    """(u_L1, p_L1, u_L2, p_L2) with QGauss(4) on the MappingQ(4) cells of `level` (an oracle ShellLevel), following
    VectorTools::integrate_difference + compute_global_error: the L1 norm of a vector integrates the sum of the absolute
    values of its components, the L2 norm is the root of the integral of the sum of their squares."""
    from .shell import LOBATTO5
    xg, wg = fe.gauss(4)
    Vm, Dm = _lagrange(np.array(LOBATTO5), xg)
    N = _tensor3(Vm)
    dN = [np.einsum("cl,bj,ai->cbalji", *[(Dm if d == r else Vm) for d in (2, 1, 0)]).reshape(64, 125) for r in range(3)]
    V2 = _tensor3(np.stack([(1 - xg) * (1 - 2 * xg), 4 * xg * (1 - xg), xg * (2 * xg - 1)], axis=1))
    V1 = _tensor3(np.stack([1 - xg, xg], axis=1))
    w = np.einsum("c,b,a->cba", wg, wg, wg).ravel()
    ks = fe.lex_nodes(5, 3)
    e = np.zeros(4)
    for cell, (t, ijk) in enumerate(level.cells):
        X = np.array([level.position(t, ijk, [LOBATTO5[k[d]] for d in range(3)]) for k in ks])
        xq = N @ X
        J = np.stack([dN[r] @ X for r in range(3)], axis=-1)            # [q][d][r]
        jxw = np.linalg.det(J) * w
        uh = np.stack([V2 @ sol_u[level.cell_dofs_u[cell, c * 27:(c + 1) * 27]] for c in range(3)], axis=-1)
        ph = V1 @ sol_p[level.cell_dofs_p[cell]]
        du = uh - velocity(xq, m)
        dp = np.abs(ph - pressure(xq, m))
        e += [(np.abs(du).sum(-1) * jxw).sum(), (dp * jxw).sum(), ((du * du).sum(-1) * jxw).sum(), (dp * dp * jxw).sum()]
    return e[0], e[1], np.sqrt(e[2]), np.sqrt(e[3])


def surface_mean_weights(level):
    """(weights[i] = integral of the i-th pressure shape function over the outer surface, area): `Pressure normalization =
    surface` on the spherical shell (helper_functions.cc:765-870: the boundary with the `top` indicator, QGauss(degree + 1)
    = 3 x 3 points per face)."""
    from .shell import LOBATTO5
    xg, wg = fe.GAUSS_X, fe.GAUSS_W
    Vm, Dm = _lagrange(np.array(LOBATTO5), xg)
    one = _lagrange(np.array(LOBATTO5), np.array([1.0]))
    ks = fe.lex_nodes(5, 3)
    wts = np.zeros(level.n_p)
    area = 0.0
    top_layer = max(ib for _, _, ib in level.trees)
    for cell, (t, ijk) in enumerate(level.cells):
        if level.trees[t][2] != top_layer or ijk[2] != level.n - 1:
            continue
        X = np.array([level.position(t, ijk, [LOBATTO5[k[d]] for d in range(3)]) for k in ks])
        for b, yq in enumerate(xg):
            for a, xq in enumerate(xg):
                # tangent vectors of the face zeta = 1
                tx = sum(X[i] * Dm[a, k[0]] * Vm[b, k[1]] * one[0][0, k[2]] for i, k in enumerate(ks))
                ty = sum(X[i] * Vm[a, k[0]] * Dm[b, k[1]] * one[0][0, k[2]] for i, k in enumerate(ks))
                da = np.linalg.norm(np.cross(tx, ty)) * wg[a] * wg[b]
                area += da
                for n, loc in enumerate(fe.lex_nodes(2, 3)):
                    if loc[2] == 1:
                        wts[level.cell_dofs_p[cell, n]] += da * (xq if loc[0] else 1 - xq) * (yq if loc[1] else 1 - yq)
    return wts, area
