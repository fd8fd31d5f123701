"""Host side of OperatorCellData: the viscosity-table fill that stays in the plugin layer.

Mirrors MatrixFreeStokesOperators::fill_active_cell_data
(/root/reference/source/simulator/solver/matrix_free_operators.cc:137-330) and the level projection of
StokesMatrixFreeHandlerLocalSmoothingImplementation::evaluate_material_model
(stokes_matrix_free_local_smoothing.cc:211-294) for the DGQ0 projection space:

* the material plugin's viscosities at the 3^dim velocity quadrature points are optionally averaged per
  cell (MaterialModel::MaterialAveraging::average, material_model/interface.cc:523-650 -- unweighted over
  the quadrature points);
* `Material averaging = none`: the active table keeps the raw values (n_q per cell, :235-246);
  otherwise the active table holds the cell's DGQ0 value (1 per cell, :312-313);
* the DGQ0 vector is the cellwise L2 projection of the (averaged) values = JxW-weighted mean
  (Utilities::project_cellwise, utilities.cc:2738-2812);
* level tables: interpolate_to_mg of the DGQ0 vector, i.e. parent = mean of its children, 1 value/cell
  (:220-226, :247-248, :274);
* min/max over the post-averaging quadrature values (:248-262) -> "Viscosity range".
"""
import numpy as np

_GW = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])

AVERAGING = ("none", "arithmetic average", "harmonic average", "geometric average",
             "harmonic average only viscosity", "geometric average only viscosity", "project to Q1", "project to Q1 only viscosity")


def quadrature_weights(dim):
    w = _GW
    for _ in range(dim - 1):
        w = np.multiply.outer(_GW, w)
    return w.ravel()


def average_viscosity(eta_q, averaging):
    op = averaging.replace(" only viscosity", "")
    if op == "none":
        return np.array(eta_q, dtype=np.float64)
    if op == "arithmetic average":
        a = eta_q.mean(axis=1)
    elif op == "harmonic average":
        with np.errstate(divide="ignore"):      # a zero value makes the harmonic mean zero (interface.cc:578-585)
            a = np.where((eta_q == 0.0).any(axis=1), 0.0, 1.0 / (1.0 / eta_q).mean(axis=1))
    elif op == "geometric average":
        a = np.prod(eta_q ** (1.0 / eta_q.shape[1]), axis=1)
    else:
        raise ValueError(f"unsupported Material averaging for the GMG path: {averaging}")
    return np.repeat(a[:, None], eta_q.shape[1], axis=1)


def _q1_at_quad(dim):
    s = np.sqrt(3.0 / 5.0)
    x = np.array([0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s])
    n1 = np.stack([1.0 - x, x], axis=1)
    phi = n1
    for _ in range(dim - 1):
        phi = np.einsum("ab,cd->acbd", n1, phi).reshape(n1.shape[0] * phi.shape[0], -1)
    return phi                                                           # [q lex][vertex lex]


def fill_cell_data_q1(dim, children, eta_q, jxw=None):
    """`project to Q1 [only viscosity]`: DGQ1 projection space (local_smoothing.cc:99-111).  Cellwise (JxW-weighted) L2
    projection of the raw values onto Q1, evaluated at the quadrature points and clamped to the raw range
    (matrix_free_operators.cc:226-232, :306-325); level transfer by interpolate_to_mg of the interpolatory DGQ1 element =
    the parent's nodal value at vertex k is child k's nodal value k; 3^dim values per cell everywhere (:274-289)."""
    v = np.asarray(eta_q, dtype=np.float64)
    phi = _q1_at_quad(dim)
    w = np.broadcast_to(quadrature_weights(dim)[None, :], v.shape) if jxw is None else np.asarray(jxw)
    eta_min, eta_max = float(v.min()), float(v.max())
    coef = np.linalg.solve(np.einsum("cq,qi,qj->cij", w, phi, phi), np.einsum("cq,qi->ci", w * v, phi)[..., None])[..., 0]
    n_levels = len(children)
    levels = [None] * n_levels
    levels[-1] = np.ascontiguousarray(np.clip(coef @ phi.T, eta_min, eta_max))
    for l in range(n_levels - 1, 0, -1):
        ch = np.asarray(children[l])
        coef = np.stack([coef[ch[:, k], k] for k in range(1 << dim)], axis=1)
        levels[l - 1] = np.ascontiguousarray(np.clip(coef @ phi.T, eta_min, eta_max))
    return levels[-1].copy(), levels, eta_min, eta_max


def fill_cell_data(dim, children, eta_q, averaging, jxw=None):
    """children[l] (l = 1..R): parent->child tables; eta_q (cells_active, 3^dim).
    Returns (active_table, [level tables 0..R], eta_min, eta_max)."""
    if averaging.startswith("project to Q1"):
        return fill_cell_data_q1(dim, children, eta_q, jxw)
    w = quadrature_weights(dim)
    out_q = average_viscosity(np.asarray(eta_q, dtype=np.float64), averaging)
    eta_min, eta_max = float(out_q.min()), float(out_q.max())
    dg0 = (out_q * w[None, :]).sum(axis=1) / w.sum()
    active = out_q if averaging == "none" else dg0[:, None]
    n_levels = len(children)
    levels = [None] * n_levels
    levels[-1] = np.ascontiguousarray(dg0[:, None])
    for l in range(n_levels - 1, 0, -1):
        levels[l - 1] = np.ascontiguousarray(levels[l][children[l], 0].mean(axis=1)[:, None])
    return np.ascontiguousarray(active), levels, eta_min, eta_max
