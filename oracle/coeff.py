"""Benchmark coefficient generators and viscosity-table semantics (oracle; test infrastructure).

* nsinker: /root/reference/benchmarks/nsinker/nsinker.cc:64-76 (formula), :196-277 (centres),
  :281-314 (inside_outside_factor), :320-331 (pressure scaling forced to 1).
* SolCx:   /root/reference/benchmarks/solcx/solcx.h:2962-2977.
* averaging: /root/reference/source/material_model/interface.cc:523-650, :850-935.
* table fill + level projection: matrix_free_operators.cc:137-330,
  stokes_matrix_free_local_smoothing.cc:211-294 (DGQ0 projection; levels get the mean of
  their children through interpolate_to_mg).
* pressure scaling: /root/reference/source/simulator/helper_functions.cc:253-348 and
  geometry_model/box.cc:252-255 (length scale = 0.01 * x-extent).
"""
import numpy as np
from . import fe

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
])  # first 24 of the 75 fixed centres (the box configs use <= 8; 8 of the 24 lie inside the shell)


def nsinker_chi(x, n_sinkers, delta=200.0, omega=0.1):
    dim = x.shape[-1]
    chi = np.ones(x.shape[:-1])
    for s in range(n_sinkers):
        dist = np.sqrt(((x - NSINKER_CENTERS[s, :dim]) ** 2).sum(-1))
        chi = chi * (1.0 - np.exp(-delta * np.maximum(0.0, dist - omega / 2.0) ** 2))
    return chi


def nsinker(x, n_sinkers=4, viscosity_ratio=1e4, sinker_density=10.0):
    """(eta, rho) at points x (..., dim)."""
    chi = nsinker_chi(x, n_sinkers)
    s = np.sqrt(viscosity_ratio)
    eta = (s - 1.0 / s) * (1.0 - chi) + 1.0 / s
    rho = sinker_density * (1.0 - chi)
    return eta, rho


def solcx(x, eta_b=1e6, background_density=0.0):
    eta = np.where(x[..., 0] < 0.5, 1.0, eta_b)
    rho = background_density - np.sin(np.pi * x[..., 1]) * np.cos(np.pi * x[..., 0])
    return eta, rho


def average(values_q, operation):
    """Cell-wise averaging of quadrature-point values (cells, nq) -> (cells, nq)
    (material_model/interface.cc:523-650; unweighted over the quadrature points)."""
    v = np.asarray(values_q, float)
    if operation == "none":
        return v.copy()
    if operation == "arithmetic":
        a = v.mean(axis=1)
    elif operation == "harmonic":
        has_zero = (v == 0.0).any(axis=1)
        with np.errstate(divide="ignore"):
            a = 1.0 / (1.0 / v).mean(axis=1)
        a = np.where(has_zero, 0.0, a)
    elif operation == "geometric":
        a = np.prod(v ** (1.0 / v.shape[1]), axis=1)
    else:
        raise ValueError(operation)
    return np.repeat(a[:, None], v.shape[1], axis=1)


def viscosity_tables(hier, eta_q, averaging):
    """Active and level viscosity tables for the DGQ0 projection path.

    eta_q: (cells_active, nq) raw plugin output on the active (= finest) mesh.
    averaging in {'none', 'harmonic', 'arithmetic', 'geometric'} (viscosity part of
    `Material averaging`).
    Returns (active_table, level_tables, eta_min, eta_max):
      active_table (cells, nq) if averaging == 'none' else (cells, 1)
      level_tables[l] (cells_l, 1) for l = 0..R
    """
    dim = hier.dim
    t = fe.tensor_tables(dim)
    if averaging == "project_q1":
        return viscosity_tables_q1(hier, eta_q)
    out_q = average(eta_q, averaging)
    eta_min, eta_max = float(out_q.min()), float(out_q.max())          # matrix_free_operators.cc:248-262
    # cellwise L2 projection onto constants = JxW-weighted mean (utilities.cc:2738-2812)
    jxw = getattr(hier.active, "jxw", None)                              # curved cells: per-cell JxW
    w = t["w"][None, :] if jxw is None else jxw
    dg0 = (out_q * w).sum(axis=1) / np.broadcast_to(w, out_q.shape).sum(axis=1)
    active = out_q.copy() if averaging == "none" else dg0[:, None].copy()
    levels = [None] * (hier.R + 1)
    levels[hier.R] = dg0[:, None].copy()
    for l in range(hier.R, 0, -1):                                      # interpolate_to_mg
        ch = hier.children(l)
        levels[l - 1] = levels[l][ch, 0].mean(axis=1)[:, None]
    return active, levels, eta_min, eta_max


def viscosity_tables_q1(hier, eta_q):
    """`Material averaging = project to Q1 [only viscosity]`: the projection space is DGQ1 (local_smoothing.cc:99-111).
    * active level: cellwise L2 projection (JxW-weighted) of the RAW quadrature-point viscosities onto Q1
      (Utilities::project_cellwise; no averaging, matrix_free_operators.cc:226-232), evaluated back at the quadrature points
      and clamped to the [min, max] of the raw values (:306-325);
    * levels: interpolate_to_mg of the DGQ1 vector -- FE_DGQ(1) is interpolatory with non-additive restriction, so a parent's
      nodal value at vertex k is child k's function value there = child k's own nodal value k -- then the same evaluation and
      clamp (local_smoothing.cc:274-289).  n_q values per cell on every level.
    Returns (active_table, level_tables, eta_min, eta_max)."""
    dim = hier.dim
    t = fe.tensor_tables(dim)
    phi = t["phi1"]                                                      # (nq, 2^dim), nodes lexicographic = vertex order
    v = np.asarray(eta_q, float)
    eta_min, eta_max = float(v.min()), float(v.max())
    jxw = getattr(hier.active, "jxw", None)
    w = np.broadcast_to(t["w"][None, :], v.shape) if jxw is None else np.asarray(jxw)
    M = np.einsum("cq,qi,qj->cij", w, phi, phi)
    b = np.einsum("cq,qi->ci", w * v, phi)
    coef = np.linalg.solve(M, b[..., None])[..., 0]                      # DGQ1 nodal values per cell
    ev = lambda c: np.clip(c @ phi.T, eta_min, eta_max)
    levels = [None] * (hier.R + 1)
    levels[hier.R] = ev(coef)
    for l in range(hier.R, 0, -1):
        ch = hier.children(l)
        coef = np.stack([coef[ch[:, k], k] for k in range(1 << dim)], axis=1)
        levels[l - 1] = ev(coef)
    return levels[hier.R].copy(), levels, eta_min, eta_max


def pressure_scaling_box(hier, eta_fn):
    """exp(volume-weighted mean of log eta at cell midpoints) / (0.01 * x-extent)."""
    lev = hier.active
    mid = (lev.cells + 0.5) * lev.h
    eta = eta_fn(mid)
    return float(np.exp(np.log(eta).mean()) / (0.01 * lev.extents[0]))


def shell_sinker_centers(inner=0.54, outer=1.0):
    """benchmarks/nsinker_spherical_shell/nsinker.cc:284-296: centres rescaled to [-1,1]^3, those outside the shell dropped."""
    p = 2.0 * NSINKER_CENTERS - 1.0
    r = np.sqrt((p * p).sum(-1))
    return p[(r >= inner) & (r <= outer)]


def nsinker_shell(x, n_sinkers=5, viscosity_ratio=1e2, sinker_density=10.0, transition_zone_jump=100.0):
    """(eta, rho) of the spherical-shell NSinker material (nsinker_spherical_shell/nsinker.cc:57-71; delta = 400,
    omega = 0.2, :204-205; viscosity times `Transition zone viscosity jump` below r = 0.89583)."""
    centers = shell_sinker_centers()
    assert n_sinkers <= len(centers)
    chi = np.ones(x.shape[:-1])
    for s in range(n_sinkers):
        dist = np.sqrt(((x - centers[s]) ** 2).sum(-1))
        chi = chi * (1.0 - np.exp(-400.0 * np.maximum(0.0, dist - 0.2 / 2.0) ** 2))
    s = np.sqrt(viscosity_ratio)
    eta = (s - 1.0 / s) * (1.0 - chi) + 1.0 / s
    eta = np.where(np.sqrt((x * x).sum(-1)) < 0.89583, eta * transition_zone_jump, eta)
    return eta, sinker_density * (1.0 - chi)


def pressure_scaling_shell(level, eta_mid, jxw_mid, inner=0.54, outer=1.0):
    """helper_functions.cc:253-348 with SphericalShell::length_scale (geometry_model/spherical_shell.cc:724-737):
    exp(JxW-weighted mean of log eta at the cell midpoints) / (1e4 / (6336000 - 3481000) * (R1 - R0))."""
    ref = np.exp((np.log(eta_mid) * jxw_mid).sum() / jxw_mid.sum())
    return float(ref / (1e4 / (6336000.0 - 3481000.0) * (outer - inner)))
