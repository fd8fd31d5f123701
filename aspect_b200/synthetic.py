"""Synthetic-mesh driver: produces, for a uniformly refined box, exactly the arrays ASPECT's handler
would pull out of deal.II in setup_dofs()
(/root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:1450-1851) and hand to
the C ABI: per-level cell->DoF maps (MatrixFree layout), constrained-DoF lists, cell sizes, the
parent->child tables and the active->level copy indices.

Numbering rule: cells in Morton order (deal.II child order, x fastest); DoFs in
first-touch order over that cell sequence and, inside a cell, FE-local order vertices -> lines ->
quads -> hex interior with the vector components of a node consecutive.  That is what
`DoFRenumbering::hierarchical` (:1488-1491) and `distribute_mg_dofs` (:1594-1596) produce on this mesh.
Everything is vectorised numpy (the refinement-7 box has 2.1 M cells / 51 M velocity DoFs).
"""
import numpy as np

_UINT = np.uint32


def _lex(npd, dim):
    g = np.indices((npd,) * dim).reshape(dim, -1)          # g[0] slowest
    return g[::-1].T.copy()                                 # (npd^dim, dim), x fastest


def _fe_local_order(degree, dim):
    """deal.II FE_Q local ordering as lexicographic multi-indices (vertices, lines, quads, hex)."""
    hi = degree
    order = [tuple(hi * ((v >> d) & 1) for d in range(dim)) for v in range(2 ** dim)]
    if degree == 2:
        m = 1
        if dim == 2:
            order += [(0, m), (hi, m), (m, 0), (m, hi), (m, m)]
        else:
            for z in (0, hi):
                order += [(0, m, z), (hi, m, z), (m, 0, z), (m, hi, z)]
            order += [(0, 0, m), (hi, 0, m), (0, hi, m), (hi, hi, m)]
            order += [(0, m, m), (hi, m, m), (m, 0, m), (m, hi, m), (m, m, 0), (m, m, hi), (m, m, m)]
    return np.array(order, dtype=np.int64)


def morton_cells(level, dim):
    n_cells = 1 << (dim * level)
    m = np.arange(n_cells, dtype=np.int64)
    out = np.zeros((n_cells, dim), dtype=np.int64)
    for b in range(level):
        for d in range(dim):
            out[:, d] |= ((m >> (dim * b + d)) & 1) << b
    return out


def _first_touch(cells, npd_minus1, local_order, lex_order, grid):
    """Number the grid nodes in first-touch order; returns (cell_nodes_lex, node_grid)."""
    dim = cells.shape[1]
    mult = grid ** np.arange(dim, dtype=np.int64)

    def keys(order):
        out = np.zeros((len(cells), len(order)), dtype=np.int64)
        for d in range(dim):
            out += (npd_minus1 * cells[:, d:d + 1] + order[None, :, d]) * mult[d]
        return out

    kfe = keys(local_order).ravel()
    uniq, first = np.unique(kfe, return_index=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    table = np.full(grid ** dim, -1, dtype=np.int64)
    table[uniq] = rank
    del kfe
    cell_nodes = table[keys(lex_order)]
    node_grid = np.zeros((len(uniq), dim), dtype=np.int64)
    for d in range(dim):
        node_grid[rank, d] = (uniq // mult[d]) % grid
    return cell_nodes, node_grid


class BoxLevel:
    def __init__(self, dim, level, extents=None):
        self.dim, self.level = dim, level
        self.n = n = 1 << level
        self.extents = np.ones(dim) if extents is None else np.asarray(extents, float)
        self.h = self.extents / n
        self.cells = morton_cells(level, dim)
        self.n_cells = len(self.cells)
        nodes_u, self.node_grid_u = _first_touch(self.cells, 2, _fe_local_order(2, dim), _lex(3, dim), 2 * n + 1)
        nodes_p, self.node_grid_p = _first_touch(self.cells, 1, _fe_local_order(1, dim), _lex(2, dim), n + 1)
        self.n_u = dim * len(self.node_grid_u)
        self.n_p = len(self.node_grid_p)
        nn = 3 ** dim
        cu = np.empty((self.n_cells, dim * nn), dtype=_UINT)
        for c in range(dim):
            cu[:, c * nn:(c + 1) * nn] = dim * nodes_u + c
        self.cell_dofs_u = cu
        self.cell_dofs_p = nodes_p.astype(_UINT)

    def constrained_u(self, bc):
        """'zero' (all components on every face), 'tangential' (normal component on each face), 'none'."""
        g, hi, dim = self.node_grid_u, 2 * self.n, self.dim
        on = (g == 0) | (g == hi)                                  # (nodes, dim)
        if bc == "zero":
            mask = np.repeat(on.any(axis=1)[:, None], dim, axis=1)
        elif bc == "tangential":
            mask = on
        elif bc == "none":
            mask = np.zeros_like(on)
        else:
            raise ValueError(bc)
        return np.flatnonzero(mask.ravel()).astype(_UINT)

    def quadrature_points(self):
        s = np.sqrt(3.0 / 5.0)
        g = np.array([0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s])
        q = _lex(3, self.dim)
        return (self.cells[:, None, :] + g[q][None, :, :]) * self.h

    def dof_coords_u(self):
        return np.repeat(self.node_grid_u * (self.h / 2.0), self.dim, axis=0)


class BoxHierarchy:
    """Levels 0..R of the uniformly refined box with one coarse cell; active mesh = level R."""

    def __init__(self, dim, refinement, extents=None):
        self.dim, self.R = dim, refinement
        self.levels = [BoxLevel(dim, l, extents) for l in range(refinement + 1)]
        self.active = self.levels[-1]

    def children(self, level):
        npar = self.levels[level - 1].n_cells
        return (np.arange(npar, dtype=np.int64)[:, None] * (1 << self.dim) + np.arange(1 << self.dim)[None, :]).astype(np.int32)
