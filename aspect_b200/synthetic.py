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


_box_lib = False


def _box_maps_lib():
    """aspect_b200/libbox_maps.so (cpp/box_maps.cc): the same numbering in plain C++ loops, used for large meshes."""
    global _box_lib
    if _box_lib is False:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbox_maps.so")
        _box_lib = None
        if os.path.exists(path) and os.environ.get("SMF_NUMPY_MAPS") != "1":
            _box_lib = ctypes.CDLL(path)
            _box_lib.box_first_touch.restype = ctypes.c_int64
    return _box_lib


def _first_touch(cells, npd_minus1, local_order, lex_order, grid):
    """Number the grid nodes in first-touch order; returns (cell_nodes_lex, node_grid).  grid: nodes per dimension."""
    dim = cells.shape[1]
    grid = np.asarray(grid, dtype=np.int64)
    lib = _box_maps_lib() if len(cells) >= 4096 else None
    if lib is not None:
        import ctypes
        c = np.ascontiguousarray(cells, dtype=np.int64)
        lo, lx = np.ascontiguousarray(local_order, dtype=np.int64), np.ascontiguousarray(lex_order, dtype=np.int64)
        cap = int(min(np.prod(grid), len(c) * len(lo)))
        cell_nodes = np.empty((len(c), len(lo)), dtype=np.int64)
        node_grid = np.empty((cap, dim), dtype=np.int64)
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        n = lib.box_first_touch(ctypes.c_int(dim), ctypes.c_int64(len(c)), vp(c), ctypes.c_int(npd_minus1), vp(grid), ctypes.c_int(len(lo)),
                                vp(lo), vp(lx), vp(cell_nodes), vp(node_grid), ctypes.c_int64(cap))
        if n < 0:
            raise RuntimeError("box_first_touch failed")
        return cell_nodes, node_grid[:n].copy()
    mult = np.concatenate([[1], np.cumprod(grid[:-1])]).astype(np.int64)

    def keys(order):
        out = np.zeros((len(cells), len(order)), dtype=np.int64)
        for d in range(dim):
            out += (npd_minus1 * cells[:, d:d + 1] + order[None, :, d]) * mult[d]
        return out

    kfe = keys(local_order).ravel()
    uniq, first = np.unique(kfe, return_index=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    del kfe
    klex = keys(lex_order)
    cell_nodes = rank[np.searchsorted(uniq, klex.ravel())].reshape(klex.shape)
    node_grid = np.zeros((len(uniq), dim), dtype=np.int64)
    for d in range(dim):
        node_grid[rank, d] = (uniq // mult[d]) % grid[d]
    return cell_nodes, node_grid


class BoxLevel:
    """One level of the (rank-local part of the) box mesh.

    The box has `repetitions` coarse cells per direction (`X/Y/Z repetitions` of the box geometry model,
    /root/reference/source/geometry_model/box.cc:66-74); every coarse cell is the root of one tree.  A rank owns a
    contiguous range of trees (p4est orders cells tree by tree along the space-filling curve) and numbers the DoFs
    of its cells locally in first-touch order; DoFs on faces shared with trees of other ranks exist on every
    rank that touches them (`sharers`), which is what the halo exchange of the C ABI works with."""

    def __init__(self, dim, level, extents=None, repetitions=None, rank=0, world=1):
        self.dim, self.level = dim, level
        self.n = n = 1 << level
        self.extents = np.ones(dim) if extents is None else np.asarray(extents, float)
        self.rep = rep = np.ones(dim, dtype=np.int64) if repetitions is None else np.asarray(repetitions, dtype=np.int64)
        self.rank, self.world = rank, world
        n_trees = int(np.prod(rep))
        assert n_trees % world == 0, "the number of coarse cells must be a multiple of the number of ranks"
        self.trees_per_rank = tpr = n_trees // world
        self.h = self.extents / (rep * n)
        m = morton_cells(level, dim)
        trees = np.arange(rank * tpr, (rank + 1) * tpr, dtype=np.int64)
        origin = np.zeros((tpr, dim), dtype=np.int64)
        t = trees.copy()
        for d in range(dim):
            origin[:, d] = (t % rep[d]) * n
            t //= rep[d]
        self.cells = (origin[:, None, :] + m[None, :, :]).reshape(-1, dim)
        self.n_cells = len(self.cells)
        self.grid_u, self.grid_p = 2 * rep * n + 1, rep * n + 1
        nodes_u, self.node_grid_u = _first_touch(self.cells, 2, _fe_local_order(2, dim), _lex(3, dim), self.grid_u)
        nodes_p, self.node_grid_p = _first_touch(self.cells, 1, _fe_local_order(1, dim), _lex(2, dim), self.grid_p)
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
        g, hi, dim = self.node_grid_u, self.grid_u - 1, self.dim
        on = (g == 0) | (g == hi[None, :])                         # (nodes, dim)
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

    # -- distributed layout --------------------------------------------------------------------------------
    def global_node_ids(self, family):
        """Partition-independent id of every local node (lexicographic index in the global node lattice)."""
        g, grid = (self.node_grid_u, self.grid_u) if family == "u" else (self.node_grid_p, self.grid_p)
        mult = np.concatenate([[1], np.cumprod(grid[:-1])]).astype(np.int64)
        return (g * mult[None, :]).sum(axis=1)

    def sharers(self, family):
        """(interface node indices, (n, 2^dim) ranks touching each of them, -1 padded / duplicates removed)."""
        g = self.node_grid_u if family == "u" else self.node_grid_p
        step = (2 if family == "u" else 1) * self.n              # lattice distance between tree planes
        dim, rep = self.dim, self.rep
        on_plane = (g % step == 0) & (g > 0) & (g < (rep * step)[None, :])
        idx = np.flatnonzero(on_plane.any(axis=1))
        gi, pl = g[idx], on_plane[idx]
        hi_tree = np.minimum(gi // step, rep[None, :] - 1)        # tree coordinate on the upper side
        ranks = np.zeros((len(idx), 1 << dim), dtype=np.int64)
        for corner in range(1 << dim):
            tree = np.zeros(len(idx), dtype=np.int64)
            mult = 1
            for d in range(dim):
                lower = ((corner >> d) & 1) == 1
                td = np.where(pl[:, d] & lower, hi_tree[:, d] - 1, hi_tree[:, d])
                tree += td * mult
                mult *= rep[d]
            ranks[:, corner] = tree // self.trees_per_rank
        ranks.sort(axis=1)
        dup = np.zeros_like(ranks, dtype=bool)
        dup[:, 1:] = ranks[:, 1:] == ranks[:, :-1]
        ranks[dup] = -1
        multi = ((ranks >= 0) & (ranks != self.rank)).any(axis=1)     # interior tree planes between two local trees are not interfaces
        return idx[multi], ranks[multi]

    def exchange_lists(self, family):
        """{neighbour rank: local node indices shared with it, ordered by global node id (identical on both sides)}."""
        idx, ranks = self.sharers(family)
        gid = self.global_node_ids(family)[idx]
        order = np.argsort(gid, kind="stable")
        idx, ranks = idx[order], ranks[order]
        out = {}
        for q in np.unique(ranks):
            if q < 0 or q == self.rank:
                continue
            out[int(q)] = idx[(ranks == q).any(axis=1)].astype(_UINT)
        return out


class BoxHierarchy:
    """Levels 0..R of the uniformly refined box (rank-local part); active mesh = level R."""

    def __init__(self, dim, refinement, extents=None, repetitions=None, rank=0, world=1):
        self.dim, self.R = dim, refinement
        self.rank, self.world = rank, world
        self.levels = [BoxLevel(dim, l, extents, repetitions, rank, world) for l in range(refinement + 1)]
        self.active = self.levels[-1]

    def children(self, level):
        npar = self.levels[level - 1].n_cells
        return (np.arange(npar, dtype=np.int64)[:, None] * (1 << self.dim) + np.arange(1 << self.dim)[None, :]).astype(np.int32)
