"""Benchmark "plugins" of the synthetic driver: material coefficients and the Stokes right-hand side.

In ASPECT these stay host C++ (material-model plugins + the FEValues assembly loop); here they are the
numpy equivalents that feed the handler:
  nsinker : /root/reference/benchmarks/nsinker/nsinker.cc:61-76, :196-314 (first 8 of the fixed centres are
            enough for the BASELINE configs), pressure scaling forced to 1 (:320-331)
  SolCx   : /root/reference/benchmarks/solcx/solcx.h:2962-2977
  rhs     : f_i = int rho g . phi_i with the 3^dim Gauss rule (source/simulator/assemblers/stokes.cc:387-396)
  pressure scaling: source/simulator/helper_functions.cc:253-348, geometry_model/box.cc:252-255
"""
import numpy as np

from .cell_data import average_viscosity, quadrature_weights

NSINKER_CENTERS = np.array([
    [2.4257829890e-01, 1.3469574514e-02, 3.8313885004e-01],
    [4.1465269048e-01, 6.7768972864e-02, 9.9312692973e-01],
    [4.8430804651e-01, 7.6533776604e-01, 3.1833815403e-02],
    [3.0935481671e-02, 9.3264044027e-01, 8.8787953411e-01],
    [5.9132973039e-01, 4.7877868473e-01, 8.3335433660e-01],
    [1.8633519681e-01, 7.3565270739e-01, 1.1505317181e-01],
    [6.9865863058e-01, 3.5560411138e-01, 6.3830000658e-01],
    [9.0821050755e-01, 2.9400041480e-01, 2.6497158886e-01],
    [3.7749399775e-01, 5.4162011554e-01, 9.2818150340e-03],
    [8.5247022139e-01, 4.6701098395e-01, 5.3607231962e-02],
    [9.7674759057e-01, 1.9675474344e-01, 8.5697294067e-01],
    [1.4421375987e-01, 8.0066218823e-01, 7.2939761948e-01],
    [9.8579064709e-01, 1.8340570954e-01, 4.9976021075e-01],
    [4.6986202126e-01, 9.7099129947e-01, 4.5077026191e-01],
    [9.5791877292e-02, 9.7408164664e-01, 3.9023506101e-01],
    [6.8067035576e-01, 2.6669318800e-02, 2.3124107450e-01],
    [4.6873909443e-01, 9.7960100555e-02, 4.1541002524e-01],
    [7.9629418710e-01, 3.1640260216e-01, 7.7853444953e-01],
    [8.2849331472e-01, 4.8714042059e-01, 3.6904878000e-01],
    [6.0284549678e-01, 2.4264360789e-02, 8.1111178631e-01],
    [3.5579259291e-01, 8.0610905439e-01, 2.7487712366e-01],
    [8.5981739865e-01, 9.5101905612e-01, 7.7727618477e-01],
    [6.8083745971e-01, 8.3518540665e-01, 9.6112961413e-01],
    [7.0542474869e-01, 7.3751226102e-02, 5.3685709440e-01],
])


def _nsinker_chi_threads(x, n_sinkers, delta, omega):
    """chi of nsinker_material for many points: the same formula evaluated chunk by chunk on all host threads."""
    import concurrent.futures as cf
    import os
    dim = x.shape[-1]
    flat = x.reshape(-1, dim)
    out = np.empty(len(flat))
    step = 1 << 16

    def work(a):
        p = flat[a:a + step]
        chi = np.ones(len(p))
        for s in range(n_sinkers):
            d = p - NSINKER_CENTERS[s, :dim]
            dist = np.sqrt(np.einsum("ij,ij->i", d, d))
            np.subtract(dist, omega / 2.0, out=dist)
            np.maximum(dist, 0.0, out=dist)
            np.multiply(dist, dist, out=dist)
            np.multiply(dist, -delta, out=dist)
            np.exp(dist, out=dist)
            np.subtract(1.0, dist, out=dist)
            chi *= dist
        out[a:a + step] = chi
    n_thr = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with cf.ThreadPoolExecutor(max_workers=n_thr) as ex:
        list(ex.map(work, range(0, len(flat), step)))
    return out.reshape(x.shape[:-1])


def nsinker_material(x, n_sinkers=8, viscosity_ratio=1e4, sinker_density=10.0, delta=200.0, omega=0.1):
    assert n_sinkers <= len(NSINKER_CENTERS)
    dim = x.shape[-1]
    if x.size > 3_000_000:                   # (numpy releases the GIL inside these kernels)
        chi = _nsinker_chi_threads(x, n_sinkers, delta, omega)
        s = np.sqrt(viscosity_ratio)
        return (s - 1.0 / s) * (1.0 - chi) + 1.0 / s, sinker_density * (1.0 - chi)
    chi = np.ones(x.shape[:-1])
    for s in range(n_sinkers):
        dist = np.sqrt(((x - NSINKER_CENTERS[s, :dim]) ** 2).sum(-1))
        chi *= 1.0 - np.exp(-delta * np.maximum(0.0, dist - omega / 2.0) ** 2)
    s = np.sqrt(viscosity_ratio)
    return (s - 1.0 / s) * (1.0 - chi) + 1.0 / s, sinker_density * (1.0 - chi)


def solcx_material(x, eta_b=1e6, background_density=0.0):
    eta = np.where(x[..., 0] < 0.5, 1.0, eta_b)
    return eta, background_density - np.sin(np.pi * x[..., 1]) * np.cos(np.pi * x[..., 0])


_S = np.sqrt(3.0 / 5.0)
_GX = np.array([0.5 - 0.5 * _S, 0.5, 0.5 + 0.5 * _S])


def _phi2_at_quad(dim):
    x = _GX
    n1 = np.stack([(1 - x) * (1 - 2 * x), 4 * x * (1 - x), x * (2 * x - 1)], axis=1)      # [q][i]
    phi = n1
    for _ in range(dim - 1):
        phi = np.einsum("ab,cd->acbd", n1, phi).reshape(n1.shape[0] * phi.shape[0], -1)   # slowest index first
    return phi                                                                             # [q lex][node lex]


def _phi1_at_quad(dim):
    x = _GX
    n1 = np.stack([1 - x, x], axis=1)
    phi = n1
    for _ in range(dim - 1):
        phi = np.einsum("ab,cd->acbd", n1, phi).reshape(n1.shape[0] * phi.shape[0], -1)
    return phi                                                                             # [q lex][node lex]


def assemble_rhs(level, rho_q, gravity):
    """Velocity right-hand side; rho_q (n_cells, 3^dim) (already averaged if `Material averaging` says so)."""
    dim = level.dim
    nn = 3 ** dim
    jxw = getattr(level, "jxw", None)                                 # cells of different sizes (adaptive meshes) carry their own JxW
    w = quadrature_weights(dim)[None, :] * float(np.prod(level.h)) if jxw is None else np.asarray(jxw)
    loc = (rho_q * w) @ _phi2_at_quad(dim)                            # (cells, nodes)
    f = np.zeros(level.n_u)
    for c in range(dim):
        if gravity[c] != 0.0:
            f += np.bincount(level.cell_dofs_u[:, c * nn:(c + 1) * nn].ravel(), weights=(gravity[c] * loc).ravel(), minlength=level.n_u)
    return f


def pressure_scaling_box(level, material):
    mid = (level.cells + 0.5) * level.h
    eta = material(mid)[0]
    return float(np.exp(np.log(eta).mean()) / (0.01 * level.extents[0]))


def nsinker_problem(hier, n_sinkers=8, viscosity_ratio=1e4, averaging="harmonic average"):
    """(viscosity_q, rhs, pressure_scaling) of benchmarks/nsinker/nsinker_gmg.prm on the given box hierarchy."""
    act = hier.active
    eta_q, rho_q = nsinker_material(act.quadrature_points(), n_sinkers, viscosity_ratio)
    if averaging != "none" and not averaging.endswith("only viscosity"):
        rho_q = average_viscosity(rho_q, averaging)          # `average` also averages the density (interface.cc:908)
    g = np.zeros(act.dim); g[-1] = -1.0
    f = assemble_rhs(act, rho_q, g)
    f[act.constrained_u("zero").astype(np.int64)] = 0.0
    return eta_q, np.concatenate([f, np.zeros(act.n_p)]), 1.0


def solcx_problem(hier, averaging="harmonic average only viscosity"):
    act = hier.active
    eta_q, rho_q = solcx_material(act.quadrature_points())
    if averaging != "none" and not averaging.endswith("only viscosity"):
        rho_q = average_viscosity(rho_q, averaging)          # (tests/no_cheap_gmg: `harmonic average` of the density too)
    g = np.zeros(act.dim); g[-1] = -1.0
    f = assemble_rhs(act, rho_q, g)
    f[act.constrained_u("tangential").astype(np.int64)] = 0.0
    return eta_q, np.concatenate([f, np.zeros(act.n_p)]), pressure_scaling_box(act, solcx_material)


# ---- benchmarks/nsinker_spherical_shell (BASELINE config 3) --------------------------------------------------------
def shell_sinker_centers(inner=0.54, outer=1.0):
    """nsinker_spherical_shell/nsinker.cc:284-296: centres rescaled from [0,1]^3 to [-1,1]^3; those outside the shell are dropped."""
    p = 2.0 * NSINKER_CENTERS - 1.0
    r = np.sqrt((p * p).sum(-1))
    return p[(r >= inner) & (r <= outer)]


def nsinker_shell_material(x, n_sinkers=5, viscosity_ratio=1e2, sinker_density=10.0, transition_zone_jump=100.0,
                           delta=400.0, omega=0.2):
    """nsinker_spherical_shell/nsinker.cc:57-71 (delta, omega: :204-205); gmg.prm: 5 sinkers, ratio 1e2, jump 100."""
    centers = shell_sinker_centers()
    assert n_sinkers <= len(centers)
    chi = np.ones(x.shape[:-1])
    for s in range(n_sinkers):
        dist = np.sqrt(((x - centers[s]) ** 2).sum(-1))
        chi *= 1.0 - np.exp(-delta * np.maximum(0.0, dist - omega / 2.0) ** 2)
    s = np.sqrt(viscosity_ratio)
    eta = (s - 1.0 / s) * (1.0 - chi) + 1.0 / s
    eta = np.where(np.sqrt((x * x).sum(-1)) < 0.89583, eta * transition_zone_jump, eta)
    return eta, sinker_density * (1.0 - chi)


def pressure_scaling_shell(level, material):
    """helper_functions.cc:253-348 (JxW-weighted log average at the cell midpoints) over SphericalShell::length_scale()
    = 1e4 / (6336000 - 3481000) * (R1 - R0) (geometry_model/spherical_shell.cc:724-737)."""
    xm, det = level.global_midpoint_geometry()           # (Utilities::MPI::sum over the ranks, helper_functions.cc:337-339)
    eta = material(xm)[0]
    ref = np.exp((np.log(eta) * det).sum() / det.sum())
    return float(ref / (1e4 / (6336000.0 - 3481000.0) * (level.R1 - level.R0)))


def nsinker_shell_problem(hier, n_sinkers=5, viscosity_ratio=1e2, averaging="harmonic average"):
    """(viscosity_q, force_q, pressure_scaling) of benchmarks/nsinker_spherical_shell/gmg.prm on a ShellHierarchy:
    force_q = rho * g at the quadrature points with `Gravity model = radial constant`, magnitude 1, pointing inwards
    (source/gravity_model/radial.cc); the right-hand side itself is integrated on the device (smf_assemble_rhs)."""
    act = hier.active
    x = act.quadrature_points()
    material = lambda y: nsinker_shell_material(y, n_sinkers, viscosity_ratio)
    eta_q, rho_q = material(x)
    if averaging != "none" and not averaging.endswith("only viscosity"):
        rho_q = average_viscosity(rho_q, averaging)
    g = -x / np.sqrt((x * x).sum(-1, keepdims=True))
    return eta_q, rho_q[..., None] * g, pressure_scaling_shell(act, material)


# ---- tests/shell_3d_gmg.prm: zero velocity at the bottom, TANGENTIAL top boundary (no-normal-flux lines on a curved boundary) ----
# data/gravity-model/prem.txt (the default profile of `Gravity model = ascii data`): depth [m], gravity [m/s^2]
PREM_GRAVITY = np.array([
    [0, 9.816], [3000, 9.822], [15000, 9.833], [24000, 9.839], [80000, 9.855], [150000, 9.878], [220000, 9.905], [310000, 9.936],
    [400000, 9.969], [500000, 9.988], [600000, 10.004], [670000, 10.014], [771000, 9.999], [971000, 9.97], [1171000, 9.947],
    [1371000, 9.933], [1571000, 9.931], [1771000, 9.947], [1971000, 9.986], [2171000, 10.054], [2371000, 10.158], [2571000, 10.31],
    [2741000, 10.484], [2771000, 10.52], [2891000, 10.682], [2971000, 10.507], [3171000, 10.046], [3371000, 9.557], [3571000, 9.041],
    [3771000, 8.502], [3971000, 7.943], [4171000, 7.365], [4371000, 6.772], [4571000, 6.167], [4771000, 5.555], [4971000, 4.941],
    [5149000, 4.4], [5171000, 4.325], [5371000, 3.62], [5571000, 2.907], [5771000, 2.186], [5971000, 1.46], [6171000, 0.731],
    [6371000, 0.0]])
SHELL_3D_RADII_M = (3481000.0, 6336000.0)


def shell_3d_gmg_problem(hier):
    """(viscosity_q, force_q, pressure_scaling) of tests/shell_3d_gmg.prm (`simple` material: viscosity 1e22, rho = 3300 (1 - 4e-5
    (T - 293)); T = 1473 K with 973 / 1973 K in the nodes of the outer / inner boundary of the Q2 temperature field; PREM
    gravity) on a ShellHierarchy whose length unit is the OUTER RADIUS (inner_radius = 3481 / 6336, outer_radius = 1: the
    driver's node keys need coordinates in [-1, 1]).  With x = s x' (s = 6,336,000 m) the system of the metric mesh divided by s
    is the system of the unit mesh with pressure scaling s * ps and force density s^2 * rho g: same Krylov iteration, same
    velocities; these scaled coefficients are what is returned."""
    s = SHELL_3D_RADII_M[1]
    act = hier.active
    assert abs(act.R1 - 1.0) < 1e-12 and abs(act.R0 - SHELL_3D_RADII_M[0] / s) < 1e-12
    x = act.quadrature_points()
    r = np.sqrt((x * x).sum(-1))
    rn = np.sqrt((act.node_x_u ** 2).sum(-1))
    T_node = np.full(len(rn), 1473.0)
    T_node[np.abs(rn - act.R1) < 1e-4] = 973.0
    T_node[np.abs(rn - act.R0) < 1e-4] = 1973.0
    cell_nodes = (np.asarray(act.cell_dofs_u[:, :27]) // 3).astype(np.int64)
    T_q = T_node[cell_nodes] @ _phi2_at_quad(3).T
    rho_q = 3300.0 * (1.0 - 4e-5 * (T_q - 293.0))
    g_mag = np.interp((1.0 - r) * s, PREM_GRAVITY[:, 0], PREM_GRAVITY[:, 1])
    force_q = (s * s) * (rho_q * g_mag)[..., None] * (-x / r[..., None])
    ps = 1e22 / (1e4 / (6336000.0 - 3481000.0) * (SHELL_3D_RADII_M[1] - SHELL_3D_RADII_M[0]))
    return np.full(x.shape[:2], 1e22), force_q, s * ps


# ---- tests/muparser_temperature_example.prm: the reference golden whose ADAPTIVE meshes follow from a `minimum refinement function`
#      alone (level 5 within 150 km of the top, else the minimum), so they can be written down without an error indicator ----
MUPARSER_BOX = dict(extents=(5e6, 1e6), repetitions=(5, 1))


def muparser_example_leaves(which):
    """Active cells (level, (i, j)) of solve `which` (1: global refinement 3; 2, 3: after one / two adaptive steps): every cell
    is coarsened to what the function and the 2:1 balance allow -- level 3 | 4, 3, 2 | 5, 4, 3, 2 from the top down (boundaries
    at 875, 750 and 500 km above the bottom)."""
    H = 1e6
    required = {1: lambda y1: 3,
                2: lambda y1: 4 if y1 > 0.875e6 else (3 if y1 > 0.75e6 else 2),
                3: lambda y1: 5 if y1 > 0.875e6 else (4 if y1 > 0.75e6 else (3 if y1 > 0.5e6 else 2))}[which]
    leaves, todo = [], [(0, (i, 0)) for i in range(MUPARSER_BOX["repetitions"][0])]
    while todo:
        l, (i, j) = todo.pop()
        if required((j + 1) * H / (1 << l)) > l:
            todo += [(l + 1, (2 * i + a, 2 * j + b)) for b in (0, 1) for a in (0, 1)]
        else:
            leaves.append((l, (i, j)))
    return sorted(leaves)


def muparser_example_problem(hier):
    """(viscosity_q, rhs before condensation, pressure_scaling) on an AdaptiveBoxHierarchy built from muparser_example_leaves:
    `simple` material (viscosity 1e21, rho = 3300 (1 - 2e-5 (T - 273))), gravity 9.81 downwards, T = the prm's half-space-cooling
    function at the Q2 nodes with the boundary temperatures (273 top, 1673 bottom) and the hanging nodes distributed."""
    from scipy.special import erfc
    act = hier.active
    xy = act.node_grid_u * (act.extents / act.hi_u)
    x, y = xy[:, 0], xy[:, 1]
    ymax, xtrm, vsub, ageop, Tm, Ts, kappa = 1.0e6, 2.2e6, 7.927e-10, 9.46e14, 1673.0, 273.0, 1e-6
    age = np.where((x > 0.0) & (x <= xtrm), x / vsub, ageop)
    T = np.where(x > 0.0, Ts + (Tm - Ts) * (1.0 - erfc((ymax - y) / (2.0 * np.sqrt(kappa * age)))), Tm)
    T[act.node_grid_u[:, 1] == act.hi_u[1]] = 273.0
    T[act.node_grid_u[:, 1] == 0] = 1673.0
    for k, (ms, ws) in act.hanging_u.items():
        T[k] = ws @ T[ms]
    T_q = T[(np.asarray(act.cell_dofs_u[:, :9]) // 2).astype(np.int64)] @ _phi2_at_quad(2).T
    rho_q = 3300.0 * (1.0 - 2e-5 * (T_q - 273.0))
    f = assemble_rhs(act, rho_q, np.array([0.0, -9.81]))
    return np.full(T_q.shape, 1e21), np.concatenate([f, np.zeros(act.n_p)]), 1e21 / (0.01 * 1e6)


# ---- benchmarks/hollow_sphere (analytic solution on the spherical shell; the reference's test hollow_sphere_gmg) ----
def _hollow_sphere_constants(m, R1=0.5, R2=1.0, gamma=1.0):
    """benchmarks/hollow_sphere/hollow_sphere.cc:72-85."""
    if m == -1:
        alpha = -gamma * (R2 ** 3 - R1 ** 3) / (R2 ** 3 * np.log(R1) - R1 ** 3 * np.log(R2))
        beta = -3 * gamma * (np.log(R2) - np.log(R1)) / (R1 ** 3 * np.log(R2) - R2 ** 3 * np.log(R1))
    else:
        alpha = gamma * (m + 1) * (R1 ** -3 - R2 ** -3) / (R1 ** (-m - 4) - R2 ** (-m - 4))
        beta = -3 * gamma * (R1 ** (m + 1) - R2 ** (m + 1)) / (R1 ** (m + 4) - R2 ** (m + 4))
    return alpha, beta


def hollow_sphere_velocity(x, m=-1, gamma=1.0):
    """`HollowSphereBoundary` = the analytic velocity (hollow_sphere.cc:56-98), spherical coordinates as in
    source/utilities.cc:697-720."""
    r = np.sqrt((x * x).sum(-1))
    phi = np.arctan2(x[..., 1], x[..., 0])
    theta = np.arccos(x[..., 2] / r)
    alpha, beta = _hollow_sphere_constants(m)
    if m == -1:
        fr = alpha / (r * r) + beta * r
        gr = -2 / (r * r) * (alpha * np.log(r) + beta / 3 * r ** 3 + gamma)
    else:
        fr = alpha / r ** (m + 3) + beta * r
        gr = -2 / (r * r) * (-alpha / (m + 1) * r ** (-m - 1) + beta / 3 * r ** 3 + gamma)
    v_r, v_t, v_p = gr * np.cos(theta), fr * np.sin(theta), fr * np.sin(theta)
    return np.stack([np.sin(theta) * np.cos(phi) * v_r + np.cos(theta) * np.cos(phi) * v_t - np.sin(phi) * v_p,
                     np.sin(theta) * np.sin(phi) * v_r + np.cos(theta) * np.sin(phi) * v_t + np.cos(phi) * v_p,
                     np.cos(theta) * v_r - np.sin(theta) * v_t], axis=-1)


def hollow_sphere_material(x, m=-1, gamma=1.0, rho_0=1000.0):
    """`HollowSphereMaterial` (hollow_sphere.cc:355-400): viscosity r^(m+1) and the density of the manufactured solution."""
    r = np.sqrt((x * x).sum(-1))
    cos_theta = x[..., 2] / r
    alpha, beta = _hollow_sphere_constants(m)
    if m == -1:
        rho = -(alpha / r ** 4 * (8 * np.log(r) - 6) + 8.0 / 3.0 * beta / r + 8 * gamma / r ** 4) * cos_theta + rho_0
    else:
        rho = -(2 * alpha * r ** -4 * (m + 3) / (m + 1) * (m - 1) - 2 * beta / 3 * (m - 1) * (m + 3) * r ** m
                - m * (m + 5) * 2 * gamma * r ** (m - 3)) * cos_theta + rho_0
    return r ** (m + 1), rho


def hollow_sphere_problem(hier, m=-1, averaging="harmonic average"):
    """(viscosity_q, force_q, pressure_scaling, boundary velocity u0 (n_u,), zero except on the boundary) of
    tests/hollow_sphere_gmg.prm on a ShellHierarchy(R, 0.5, 1.0, n_coarse=12): radial gravity of magnitude 1, velocity
    prescribed on both boundaries from the analytic solution (`Prescribed velocity boundary indicators`)."""
    act = hier.active
    x = act.quadrature_points()
    material = lambda y: hollow_sphere_material(y, m)
    eta_q, rho_q = material(x)
    if averaging != "none" and not averaging.endswith("only viscosity"):
        rho_q = average_viscosity(rho_q, averaging)
    g = -x / np.sqrt((x * x).sum(-1, keepdims=True))
    u0 = np.zeros(act.n_u)
    cons = act.constrained_u("zero").astype(np.int64)
    u0[cons] = hollow_sphere_velocity(act.node_x_u, m).ravel()[cons]
    return eta_q, rho_q[..., None] * g, pressure_scaling_shell(act, material), u0


# ---- cookbooks/convection_box_3d (BASELINE config 4): the Stokes side of a time step ------------------------------------
def convection_box_temperature(x, perturbation=0.01):
    """Initial temperature of convection_box_3d.prm:49-57, (1 - z) - p cos(pi x) sin(pi z) y^3; a time loop changes p here."""
    return (1.0 - x[..., 2]) - perturbation * np.cos(np.pi * x[..., 0]) * np.sin(np.pi * x[..., 2]) * x[..., 1] ** 3


def convection_box_problem(hier, perturbation=0.01, rayleigh=1e6):
    """(viscosity_q, rhs, pressure_scaling) of the Stokes solve of one time step of cookbooks/convection_box_3d on a box
    hierarchy: viscosity 1, tangential velocity on all six faces, buoyancy Ra * T in z (density 1 - 1e-10 T under a gravity
    of 1e16, convection_box_3d.prm:91-110, written in the reduced-pressure form: the hydrostatic part of the weight is a
    gradient that the pressure absorbs).  `Material averaging` is none (the cookbook's default)."""
    act = hier.active
    x = act.quadrature_points()
    g = np.zeros(act.dim); g[-1] = 1.0
    f = assemble_rhs(act, rayleigh * convection_box_temperature(x, perturbation), g)
    f[act.constrained_u("tangential").astype(np.int64)] = 0.0
    return np.ones(x.shape[:-1]), np.concatenate([f, np.zeros(act.n_p)]), 1.0 / (0.01 * float(act.extents[0]))
