"""Adaptively refined box meshes for the synthetic driver: the arrays ASPECT's handler pulls out of deal.II / p4est on an
adaptively refined `parallel::distributed::Triangulation` of a one-tree box and hands to the C ABI --

* the ACTIVE mesh: leaves of the tree in z-order, 2:1 balanced over faces, edges and corners (`P4EST_CONNECT_FULL`), DoFs by
  first touch (`distribute_dofs` + `DoFRenumbering::hierarchical`, stokes_matrix_free_local_smoothing.cc:1488-1491), the
  hanging-node lines of `DoFTools::make_hanging_node_constraints` (:1508, :1580) in closed form, per-cell sizes;
* the LEVEL meshes of local smoothing: level l = all cells of level l, `distribute_mg_dofs` numbering (:1594-1596), boundary
  indices and refinement-edge indices of `MGConstrainedDoFs` (:1733-1773), parent -> child tables with -1 where a parent is
  not refined, and the `MGTransferMatrixFree` copy indices between active and level DoFs;
* the DGQ0 level viscosity tables of `interpolate_to_mg` on such a hierarchy (:220-226): a level cell carries its own value
  where it is active, else the mean of its children.

Vectorised numpy over integer lattice keys; tests/test_adaptive_driver.py checks the arrays against an independent,
dictionary-based restatement kept with the test infrastructure.
"""
import itertools

import numpy as np

from .synthetic import _fe_local_order, _lex

_UINT = np.uint32


def _morton(ijk, nbits):
    dim = ijk.shape[1]
    m = np.zeros(len(ijk), dtype=np.int64)
    for b in range(nbits):
        for d in range(dim):
            m |= ((ijk[:, d] >> b) & 1) << (dim * b + d)
    return m


def _tree_code(ijk, level, reps):
    """Sort key of level-`level` cell positions in p4est order: tree by tree (trees of a subdivided box: x fastest), Morton order
    inside a tree."""
    dim = ijk.shape[1]
    nbits = max(level, 1)
    tree = ijk >> level
    t = np.zeros(len(ijk), dtype=np.int64)
    mult = 1
    for d in range(dim):
        t += tree[:, d] * mult
        mult *= int(reps[d])
    return (t << (dim * nbits)) | _morton(ijk - (tree << level), nbits)


def _first_touch(keys_fe, keys_lex):
    """Number the distinct keys in the order of their first occurrence in keys_fe (cells x local order, row-major);
    returns (cell_nodes for keys_lex, sorted unique keys, number of every sorted key)."""
    flat = keys_fe.ravel()
    uniq, first = np.unique(flat, return_index=True)
    num = np.empty(len(uniq), dtype=np.int64)
    num[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    return num[np.searchsorted(uniq, keys_lex.ravel())].reshape(keys_lex.shape), uniq, num


class _Mesh:
    """Cells of one mesh (the active one or a level) with their Q2 / Q1 numbering.  Lattice keys: node position in units of
    the finest cell size / 2 (Q2) or / 1 (Q1), packed x-fastest into one int64."""

    def __init__(self, dim, levels, ijk, Lmax, extents, reps=None):
        self.dim, self.Lmax = dim, Lmax
        self.cell_level, self.cells = levels, ijk
        self.n_cells = len(ijk)
        self.extents = extents
        self.reps = np.ones(dim, dtype=np.int64) if reps is None else np.asarray(reps, dtype=np.int64)
        self.hi_u, self.hi_p = self.reps * (2 << Lmax), self.reps * (1 << Lmax)      # largest lattice coordinate per direction
        self.G2, self.G1 = int(self.hi_u.max()) + 1, int(self.hi_p.max()) + 1        # packing bases of the lattice keys
        shift = (Lmax - levels)[:, None]

        def keys(order, per, G):
            pos = ((per * ijk[:, None, :] + order[None, :, :]) << shift[:, :, None])       # (cells, nodes, dim)
            return sum(pos[:, :, d] * G ** d for d in range(dim))
        nu, self.keys_u, self.num_u = _first_touch(keys(_fe_local_order(2, dim), 2, self.G2), keys(_lex(3, dim), 2, self.G2))
        npn, self.keys_p, self.num_p = _first_touch(keys(_fe_local_order(1, dim), 1, self.G1), keys(_lex(2, dim), 1, self.G1))
        nn = 3 ** dim
        cu = np.empty((self.n_cells, dim * nn), dtype=_UINT)
        for c in range(dim):
            cu[:, c * nn:(c + 1) * nn] = dim * nu + c
        self.cell_dofs_u, self.cell_dofs_p = cu, npn.astype(_UINT)
        self.n_u, self.n_p = dim * len(self.keys_u), len(self.keys_p)
        # lattice coordinates of every node, in node-number order
        self.node_grid_u = np.empty((len(self.keys_u), dim), dtype=np.int64)
        self.node_grid_p = np.empty((len(self.keys_p), dim), dtype=np.int64)
        for d in range(dim):
            self.node_grid_u[self.num_u, d] = (self.keys_u // self.G2 ** d) % self.G2
            self.node_grid_p[self.num_p, d] = (self.keys_p // self.G1 ** d) % self.G1

    def node_of_key(self, keys, family):
        """node number of lattice keys (-1 where no node of this mesh sits)."""
        uniq, num = (self.keys_u, self.num_u) if family == "u" else (self.keys_p, self.num_p)
        pos = np.clip(np.searchsorted(uniq, keys), 0, len(uniq) - 1)
        return np.where(uniq[pos] == keys, num[pos], -1)

    def constrained_u(self, bc):
        """Dirichlet DoFs on the box boundary: 'zero' | 'tangential' | 'none' (as synthetic.BoxLevel)."""
        g, dim = self.node_grid_u, self.dim
        on = (g == 0) | (g == self.hi_u[None, :])
        if bc == "zero":
            mask = np.repeat(on.any(axis=1)[:, None], dim, axis=1)
        elif bc == "tangential":
            mask = on
        elif bc == "none":
            mask = np.zeros_like(on)
        else:
            raise ValueError(bc)
        return np.flatnonzero(mask.ravel()).astype(_UINT)

    def dof_coords_u(self):
        return np.repeat(self.node_grid_u * (self.extents / self.hi_u), self.dim, axis=0)


class AdaptiveLevel(_Mesh):
    """Level l of local smoothing: uniform cell size, boundary + refinement-edge indices."""

    def __init__(self, dim, level, ijk, Lmax, extents, reps=None):
        super().__init__(dim, np.full(len(ijk), level, dtype=np.int64), ijk, Lmax, extents, reps)
        self.level = level
        self.h = extents / (self.reps << level)
        self.codes = _tree_code(ijk, level, self.reps)             # sorted (cells are in tree-major Morton order)
        self.edge_u = self._edge(self.node_grid_u, 2, True)
        self.edge_p = self._edge(self.node_grid_p, 1, False)

    def cell_index(self, ijk):
        """index of level-cell positions ijk (n, dim) in this level's mesh, -1 where the level has no such cell."""
        n = (self.reps << self.level)[None, :]
        inside = ((ijk >= 0) & (ijk < n)).all(axis=1)
        code = _tree_code(np.clip(ijk, 0, n - 1), self.level, self.reps)
        pos = np.clip(np.searchsorted(self.codes, code), 0, len(self.codes) - 1)
        return np.where(inside & (self.codes[pos] == code), pos, -1)

    def _edge(self, grid, per, vector):
        """DoFs of the nodes that touch a level-l cell position inside the box which is not part of the level mesh."""
        span = per << (self.Lmax - self.level)
        n = (self.reps << self.level)[None, :]
        edge = np.zeros(len(grid), dtype=bool)
        for off in itertools.product((0, 1), repeat=self.dim):
            q = (grid - np.array(off)) // span
            q = np.clip(q, 0, n - 1)                                 # positions outside the box fold onto the boundary cell
            edge |= self.cell_index(q) < 0
        idx = np.flatnonzero(edge)
        if not vector:
            return idx.astype(_UINT)
        return (self.dim * idx[:, None] + np.arange(self.dim)[None, :]).ravel().astype(_UINT)


class AdaptiveActive(_Mesh):
    """The active mesh: cells of different sizes, hanging-node lines."""

    def __init__(self, dim, levels, ijk, Lmax, extents, reps=None):
        super().__init__(dim, levels, ijk, Lmax, extents, reps)
        self.h_cell = (extents / self.reps)[None, :] / (1 << levels)[:, None]       # (cells, dim)
        self.h = extents / (self.reps << Lmax)                                      # (unused: the cells go in with geometry = 1)
        self.hanging_u = self._hanging("u", 3, lambda x: np.stack([(1 - x) * (1 - 2 * x), 4 * x * (1 - x), x * (2 * x - 1)], -1))
        self.hanging_p = self._hanging("p", 2, lambda x: np.stack([1 - x, x], -1))

    # -- geometry handed to the C ABI as `geometry = 1` (cells of different sizes share one kernel launch) ---------------
    def quadrature_points(self):
        s = np.sqrt(3.0 / 5.0)
        g = np.array([0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s])
        return (self.cells[:, None, :] + g[_lex(3, self.dim)][None, :, :]) * self.h_cell[:, None, :]

    @property
    def jinv(self):
        nq, dim = 3 ** self.dim, self.dim
        out = np.zeros((self.n_cells, nq, dim, dim))
        for d in range(dim):
            out[:, :, d, d] = (1.0 / self.h_cell[:, d])[:, None]
        return out

    @property
    def jxw(self):
        from .cell_data import quadrature_weights
        return self.h_cell.prod(axis=1)[:, None] * quadrature_weights(self.dim)[None, :]

    # -- hanging nodes: coarse-cell centric -----------------------------------------------------------------------------
    def _hanging(self, family, npd, shape):
        """{hanging node: (master nodes, weights)}.  For every active cell C: the points of the half-spacing lattice on C's
        boundary that are not nodes of C but are nodes of the mesh belong to finer neighbours and hang on C: their value is
        C's interpolant there (tensor product of the 1D shape functions at multiples of 1 / (2 (npd - 1)))."""
        dim, per = self.dim, npd - 1
        G = self.G2 if family == "u" else self.G1
        nsub = 2 * per + 1                                            # half-lattice points per direction on a cell: 5 (Q2) / 3 (Q1)
        sub = _lex(nsub, dim)                                         # (nsub^dim, dim), x fastest
        on_bnd = ((sub == 0) | (sub == nsub - 1)).any(axis=1)
        odd = (sub % 2 == 1).any(axis=1)
        cand = sub[on_bnd & odd]
        # 1D weights of the candidates with respect to the cell's npd nodes
        w1 = shape(np.arange(nsub) / (nsub - 1.0))                    # (nsub, npd)
        lexn = _lex(npd, dim)
        W = np.ones((len(cand), len(lexn)))
        for d in range(dim):
            W *= w1[cand[:, d]][:, lexn[:, d]]
        lines = {}
        coarse = np.flatnonzero(self.cell_level < self.Lmax)
        if len(coarse) == 0:
            return lines
        sp = 1 << (self.Lmax - self.cell_level[coarse] - 1)          # spacing of C's half lattice in lattice units (both families)
        base = (per * self.cells[coarse]) << (self.Lmax - self.cell_level[coarse])[:, None]       # lattice position of C's corner
        pos = base[:, None, :] + cand[None, :, :] * sp[:, None, None]                     # (coarse, cand, dim)
        keys = sum(pos[:, :, d] * G ** d for d in range(dim))
        node = self.node_of_key(keys.ravel(), family).reshape(keys.shape)
        cell_nodes = (self.cell_dofs_u[:, :3 ** dim].astype(np.int64) // dim) if family == "u" else self.cell_dofs_p.astype(np.int64)
        ci, ki = np.nonzero(node >= 0)
        for c, k in zip(ci, ki):
            n = int(node[c, k])
            if n in lines:
                continue
            w = W[k]
            nz = np.abs(w) > 1e-14
            lines[n] = (cell_nodes[coarse[c]][nz].copy(), w[nz].copy())
        # close the lines (masters that hang themselves), as AffineConstraints::close() does
        changed = True
        while changed:
            changed = False
            for n, (ms, ws) in list(lines.items()):
                if any(int(m) in lines for m in ms):
                    acc = {}
                    for m, w in zip(ms, ws):
                        if int(m) in lines:
                            for m2, w2 in zip(*lines[int(m)]):
                                acc[int(m2)] = acc.get(int(m2), 0.0) + w * w2
                        else:
                            acc[int(m)] = acc.get(int(m), 0.0) + w
                    keys_sorted = sorted(acc)
                    lines[n] = (np.array(keys_sorted, dtype=np.int64), np.array([acc[k] for k in keys_sorted]))
                    changed = True
        return lines

    def constraint_lines(self, family, bc="zero"):
        """CSR arrays of the closed hanging-node lines per DoF for smf_set_constraints: (constrained_dof, row_start, master_dof,
        weight).  Velocity: one line per component; a hanging DoF on the Dirichlet boundary is a Dirichlet DoF (no line), and
        Dirichlet masters are dropped (they read zero and receive nothing)."""
        dim = self.dim
        if family == "p":
            ks = sorted(self.hanging_p)
            rows = [(k, self.hanging_p[k][0], self.hanging_p[k][1]) for k in ks]
        else:
            dirichlet = np.zeros(self.n_u, dtype=bool)
            dirichlet[self.constrained_u(bc).astype(np.int64)] = True
            rows = []
            for k in sorted(self.hanging_u):
                ms, ws = self.hanging_u[k]
                for c in range(dim):
                    if dirichlet[dim * k + c]:
                        continue
                    md = dim * ms + c
                    keep = ~dirichlet[md]
                    rows.append((dim * k + c, md[keep], ws[keep]))
        dof = np.array([r[0] for r in rows], dtype=_UINT)
        start = np.concatenate([[0], np.cumsum([len(r[1]) for r in rows])]).astype(_UINT)
        master = np.concatenate([r[1] for r in rows]).astype(_UINT) if rows else np.zeros(0, dtype=_UINT)
        weight = np.concatenate([r[2] for r in rows]).astype(np.float64) if rows else np.zeros(0)
        return dof, start, master, weight


class AdaptiveBoxHierarchy:
    """Active mesh + local-smoothing levels of an adaptively refined box (one coarse cell, or `repetitions` of them)."""

    adaptive = True

    def __init__(self, dim, leaves, extents=None, repetitions=None):
        """leaves: iterable of (level, (i, j[, k])), 0 <= i < repetitions[0] * 2^level etc.; extents: of the whole box;
        repetitions: coarse cells (trees) per direction (`X / Y / Z repetitions`), default one."""
        self.dim = dim
        self.extents = np.ones(dim) if extents is None else np.asarray(extents, float)
        self.reps = reps = np.ones(dim, dtype=np.int64) if repetitions is None else np.asarray(repetitions, dtype=np.int64)
        lv = np.array([l for l, _ in leaves], dtype=np.int64)
        ijk = np.array([c for _, c in leaves], dtype=np.int64).reshape(-1, dim)
        self.Lmax = Lmax = int(lv.max())
        order = np.argsort(_tree_code(ijk << (Lmax - lv)[:, None], Lmax, reps), kind="stable")
        lv, ijk = lv[order], ijk[order]
        self.active = AdaptiveActive(dim, lv, ijk, Lmax, self.extents, reps)
        self.R = Lmax
        # level meshes: every leaf and all its ancestors
        self.levels = []
        cur_l, cur = lv.copy(), ijk.copy()
        per_level = [None] * (Lmax + 1)
        for l in range(Lmax, -1, -1):
            here = cur[cur_l == l]
            here = np.unique(here, axis=0) if len(here) else here
            per_level[l] = here[np.argsort(_tree_code(here, l, reps), kind="stable")]
            up = cur_l == l
            cur = np.where(up[:, None], cur >> 1, cur)
            cur_l = np.where(up, l - 1, cur_l)
        self.levels = [AdaptiveLevel(dim, l, per_level[l], Lmax, self.extents, reps) for l in range(Lmax + 1)]

    @classmethod
    def refined(cls, dim, base_level, marker, n_adaptive, extents=None):
        """Uniform refinement to base_level, then n_adaptive rounds: refine the active cells with marker(center, level) true
        (centers in the unit box) and whatever the 2:1 balance over faces, edges and corners then requires."""
        leaves = {(base_level, ijk) for ijk in itertools.product(range(1 << base_level), repeat=dim)}
        for _ in range(n_adaptive):
            marked = [c for c in leaves if marker(np.array([(i + 0.5) / (1 << c[0]) for i in c[1]]), c[0])]
            queue = list(marked)
            while queue:
                c = queue.pop()
                if c not in leaves:
                    continue
                leaves.remove(c)
                l, ijk = c
                leaves.update((l + 1, tuple(2 * ijk[d] + ((k >> d) & 1) for d in range(dim))) for k in range(1 << dim))
                n = 1 << l
                for off in itertools.product((-1, 0, 1), repeat=dim):          # neighbours one level coarser must be refined, too
                    q = tuple(ijk[d] + off[d] for d in range(dim))
                    if any(off) and all(0 <= q[d] < n for d in range(dim)):
                        anc = (l - 1, tuple(i // 2 for i in q))
                        if anc in leaves:
                            queue.append(anc)
        return cls(dim, sorted(leaves), extents)

    # -- tables between the meshes ---------------------------------------------------------------------------------------
    def children(self, level):
        """(n_cells(level - 1), 2^dim) cells of `level` under every cell of `level - 1`, -1 where the parent is not refined."""
        par, F, dim = self.levels[level - 1], self.levels[level], self.dim
        out = np.empty((par.n_cells, 1 << dim), dtype=np.int32)
        for k in range(1 << dim):
            off = np.array([(k >> d) & 1 for d in range(dim)])
            out[:, k] = F.cell_index(2 * par.cells + off)
        return out

    def active_cells_on_level(self, level):
        """(indices of the active cells of level `level`, their indices in the level mesh)."""
        act = self.active
        ia = np.flatnonzero(act.cell_level == level)
        return ia, self.levels[level].cell_index(act.cells[ia])

    def copy_indices(self, level):
        """MGTransfer copy indices of one level: (active_u, level_u, active_p, level_p) -- the DoFs of the active cells of this
        level that are not on the level's refinement edge (those are copied one level down, where they are ordinary DoFs)."""
        ia, il = self.active_cells_on_level(level)
        act, L = self.active, self.levels[level]
        out = []
        for cd_a, cd_l, edge, n_l in ((act.cell_dofs_u, L.cell_dofs_u, L.edge_u, L.n_u), (act.cell_dofs_p, L.cell_dofs_p, L.edge_p, L.n_p)):
            a, b = cd_a[ia].ravel().astype(np.int64), cd_l[il].ravel().astype(np.int64)
            on_edge = np.zeros(n_l, dtype=bool)
            on_edge[edge.astype(np.int64)] = True
            keep = ~on_edge[b]
            a, b = a[keep], b[keep]
            ua, first = np.unique(a, return_index=True)
            out += [ua.astype(_UINT), b[first].astype(_UINT)]
        return tuple(out)

    def level_viscosity(self, eta_active):
        """DGQ0 value of every level cell from the active cells' values (n_active,): own value where active, else the mean of
        the children (interpolate_to_mg)."""
        out = [None] * (self.Lmax + 1)
        for l in range(self.Lmax, -1, -1):
            L = self.levels[l]
            v = np.zeros(L.n_cells)
            if l < self.Lmax:
                ch = self.children(l + 1)
                ref = ch[:, 0] >= 0
                v[ref] = out[l + 1][ch[ref]].mean(axis=1)
            ia, il = self.active_cells_on_level(l)
            v[il] = np.asarray(eta_active)[ia]
            out[l] = v
        return out
