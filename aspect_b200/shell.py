"""Synthetic spherical-shell driver (BASELINE config 3, SURVEY 8(d) "S-shell"): the arrays ASPECT's handler would pull out
of deal.II for `Geometry model = spherical shell` in 3D -- per-level cell->DoF maps, constrained DoFs, parent->child
tables and, because the cells are curved, the per-quadrature-point inverse Jacobians and JxW values that
`MatrixFree::get_mapping_info()` holds.

Coarse mesh = what `GridGenerator::hyper_shell(..., 96, true)` produces
(/root/reference/source/geometry_model/spherical_shell.cc:361-373): the 12 rhombic faces of a rhombic dodecahedron
(14 vertices: the 8 cube corners and the 6 axis points, projected onto the sphere), each split 2 x 2, times 2 radial
layers = 96 hexahedra with 150 vertices (the reference golden `tests/spherical_nsinker_gmg_gc`: "coarse size S: 150").
Geometry = what deal.II produces on a SphericalManifold (restated from its documented algorithms; pinned by the analytic
errors of the reference golden tests/hollow_sphere_gmg, which this geometry reproduces to all printed digits, and by the
`Viscosity range: 0.1 - 433.496` line of tests/spherical_nsinker_gmg_gc):
  * refinement: uniform in the radius; on the sphere a new edge vertex is the normalised mean of the edge's end points, a
    new face vertex the weighted spherical average of the face's 4 vertices (-1/4) and 4 new edge vertices (+1/2)
    (`quad->center(true, true)` in Triangulation::execute_refinement);
  * MappingQ(4) (source/simulator/core.cc:118-119): the 5^3 Gauss-Lobatto support points of a cell are the weighted
    spherical averages of the cell's own vertices with the Q1 shape functions as weights (direction: minimiser of the
    weighted squared geodesic distances, SphericalManifold::get_new_points; radius: weighted mean), and the cell map is
    their degree-4 Lagrange interpolant.
Every coarse cell is the root of one tree; level-l cells are its 8^l descendants in Morton order (deal.II child order),
trees in coarse-cell order -- the ORDER of the coarse cells (hence the DoF numbering) is this driver's own choice.
Boundary ids: 0 = inner, 1 = outer (`Zero velocity boundary indicators = top, bottom` constrains both).

DoF numbering: first touch over that cell sequence, FE-local order inside a cell, exactly as for the
box driver (synthetic.py); nodes shared between trees are identified by their position.
"""
import numpy as np

from .synthetic import _fe_local_order, _lex, morton_cells

_UINT = np.uint32
_S = np.sqrt(3.0 / 5.0)
GAUSS3 = np.array([0.5 - 0.5 * _S, 0.5, 0.5 + 0.5 * _S])
GAUSS3_W = np.array([5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0])
_L = np.sqrt(3.0 / 7.0)
LOBATTO5 = np.array([0.0, 0.5 - 0.5 * _L, 0.5, 0.5 + 0.5 * _L, 1.0])


def coarse_shell(n_cells=96):
    """Corner directions (n_cells, 4, 3) (lexicographic: v00, v10, v01, v11, unit vectors, right-handed with the outward radial
    direction) of the surface quad of every coarse cell and its radial interval as indices into (R0, Rm, R1).
    n_cells = 96 (`GridGenerator::hyper_shell(.., 96, true)`, the default of the 3D spherical shell): every rhombic face split
    2 x 2, two radial layers; n_cells = 12 (`Cells along circumference = 12`): the 12 rhombic faces, one layer [R0, R1]."""
    assert n_cells in (12, 96)
    quads = []
    for k in range(3):                                   # the cube edge runs along axis k
        i, j = (k + 1) % 3, (k + 2) % 3
        for sj in (-1.0, 1.0):
            for si in (-1.0, 1.0):
                c1 = np.zeros(3); c1[i], c1[j], c1[k] = si, sj, -1.0
                c2 = c1.copy(); c2[k] = 1.0
                a1 = np.zeros(3); a1[i] = si
                a2 = np.zeros(3); a2[j] = sj
                v = [a1, c1 / np.sqrt(3.0), c2 / np.sqrt(3.0), a2]          # v00, v10, v01, v11 (cyclic: a1, c1, a2, c2)
                if np.dot(np.cross(v[1] - v[0], v[2] - v[0]), v[0] + v[3]) < 0.0:
                    v[1], v[2] = v[2], v[1]
                quads.append(np.array(v))
    if n_cells == 12:
        return np.array(quads), np.zeros(12, dtype=np.int64), np.full(12, 2, dtype=np.int64)
    dirs, layer = [], []
    for v in quads:
        # 3 x 3 lattice of the face, projected onto the sphere
        g = np.zeros((3, 3, 3))
        for b in range(3):
            for a in range(3):
                x, y = a / 2.0, b / 2.0
                p = (1 - x) * (1 - y) * v[0] + x * (1 - y) * v[1] + (1 - x) * y * v[2] + x * y * v[3]
                g[b, a] = p / np.linalg.norm(p)
        for r in range(2):
            for sy in range(2):
                for sx in range(2):
                    dirs.append(np.array([g[sy, sx], g[sy, sx + 1], g[sy + 1, sx], g[sy + 1, sx + 1]]))
                    layer.append(r)
    layer = np.array(layer, dtype=np.int64)
    return np.array(dirs), layer, layer + 1


def sph_avg(P, W, max_iter=200):
    """Weighted spherical average of unit vectors, vectorised: P (..., m, 3), W (..., m) with sum 1 -> (..., 3): the point x
    on the sphere with sum_i W_i log_x(P_i) = 0 (minimiser of the weighted squared geodesic distances), by the fixed-point
    iteration x <- exp_x(sum_i W_i log_x(P_i)) from the normalised linear average."""
    x = (W[..., None] * P).sum(-2)
    x /= np.sqrt((x * x).sum(-1, keepdims=True))
    for _ in range(max_iter):
        c = (P * x[..., None, :]).sum(-1)
        t = P - c[..., None] * x[..., None, :]
        nt = np.sqrt((t * t).sum(-1))
        f = np.where(nt > 1e-300, np.arctan2(nt, c) / np.maximum(nt, 1e-300), 0.0)
        v = ((W * f)[..., None] * t).sum(-2)
        nv = np.sqrt((v * v).sum(-1, keepdims=True))
        if nv.max() < 1e-15:
            break
        x = np.cos(nv) * x + np.where(nv > 1e-300, np.sin(nv) / np.maximum(nv, 1e-300), 1.0) * v
        x /= np.sqrt((x * x).sum(-1, keepdims=True))
    return x


def refine_lattice(D):
    """Surface vertex directions of the children: D (trees, n+1, n+1, 3) -> (trees, 2n+1, 2n+1, 3)."""
    T, n = D.shape[0], D.shape[1] - 1
    unit = lambda v: v / np.sqrt((v * v).sum(-1, keepdims=True))
    F = np.zeros((T, 2 * n + 1, 2 * n + 1, 3))
    F[:, ::2, ::2] = D
    F[:, ::2, 1::2] = unit(D[:, :, :-1] + D[:, :, 1:])
    F[:, 1::2, ::2] = unit(D[:, :-1, :] + D[:, 1:, :])
    P = np.stack([D[:, :-1, :-1], D[:, :-1, 1:], D[:, 1:, :-1], D[:, 1:, 1:],
                  F[:, 0:-1:2, 1::2], F[:, 2::2, 1::2], F[:, 1::2, 0:-1:2], F[:, 1::2, 2::2]], axis=-2)
    W = np.broadcast_to(np.array([-0.25] * 4 + [0.5] * 4), P.shape[:-1])
    F[:, 1::2, 1::2] = sph_avg(P, W)
    return F


def _lagrange_tables(nodes, x):
    """values and derivatives of the 1D Lagrange basis on `nodes` at the points x: ([len(x)][len(nodes)]) each."""
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
                    if c != a and c != b:
                        t *= (x - nodes[c]) / (nodes[a] - nodes[c])
                D[:, a] += t
    return V, D


def mapping_tables(degree=4, points=GAUSS3):
    """Tensor Lagrange basis of MappingQ(degree) at the 3^3 tensor points: N[q][k], dN[q][k][r] (q, k lexicographic)."""
    nodes = LOBATTO5 if degree == 4 else np.linspace(0.0, 1.0, degree + 1)
    V, D = _lagrange_tables(nodes, points)
    nq, nk = len(points), len(nodes)
    N = np.einsum("ck,bj,ai->cbakji", V, V, V).reshape(nq ** 3, nk ** 3)
    dN = np.stack([np.einsum("ck,bj,ai->cbakji", V, V, D).reshape(nq ** 3, nk ** 3),
                   np.einsum("ck,bj,ai->cbakji", V, D, V).reshape(nq ** 3, nk ** 3),
                   np.einsum("ck,bj,ai->cbakji", D, V, V).reshape(nq ** 3, nk ** 3)], axis=-1)
    return nodes, N, dN


class ShellLevel:
    """One level of the (rank-local part of the) uniformly refined 96-tree spherical shell.

    A rank owns a contiguous range of trees (p4est orders the cells tree by tree) and numbers the DoFs of its cells locally
    in first-touch order; nodes on faces shared with trees of other ranks exist on every rank that touches them, which is
    what the halo exchange of the C ABI works with (`exchange_lists`)."""
    dim = 3
    KEY_BITS = 21          # per-coordinate field of the position key; quantum 2^-19

    def __init__(self, level, inner_radius=0.54, outer_radius=1.0, mapping_degree=4, rank=0, world=1, n_coarse=96):
        self.level = level
        self.n_coarse = n_coarse
        self.n = n = 1 << level
        self.R0, self.R1 = float(inner_radius), float(outer_radius)
        self.mapping_degree = mapping_degree
        self.rank, self.world = rank, world
        dirs, ia, ib = coarse_shell(n_coarse)
        self.n_trees = len(dirs)
        assert self.n_trees % world == 0, "the number of coarse cells must be a multiple of the number of ranks"
        self.trees_per_rank = tpr = self.n_trees // world
        m = self._morton = morton_cells(level, 3)
        self.cells = np.tile(m, (tpr, 1))                                 # integer coordinates inside the tree
        self.tree = np.repeat(np.arange(rank * tpr, (rank + 1) * tpr), len(m))
        self.n_cells = len(self.cells)
        radii = np.array([self.R0, 0.5 * (self.R0 + self.R1), self.R1])
        self._ra, self._rb = radii[ia], radii[ib]
        # surface vertex lattice of every tree at this level: _lattice[t, j, i] = direction of vertex (i, j)
        D = np.stack([np.stack([dirs[:, 0], dirs[:, 1]], axis=1), np.stack([dirs[:, 2], dirs[:, 3]], axis=1)], axis=1)
        for _ in range(level):
            D = refine_lattice(D)
        self._lattice = D
        self.h = np.full(3, 1.0 / n)                                      # (reference-cell size inside the tree; unused by the curved kernels)
        self._keys = {}
        nodes_u, self.node_x_u = self._first_touch("u", 2, _fe_local_order(2, 3), _lex(3, 3))
        nodes_p, self.node_x_p = self._first_touch("p", 1, _fe_local_order(1, 3), _lex(2, 3))
        n_surface = (48 if n_coarse == 96 else 12) * 4 ** level          # surface quads; radial cells: 2 n (96) or n (12)
        nr = (2 if n_coarse == 96 else 1) * n
        self.n_global_nodes = {"u": (4 * n_surface + 2) * (2 * nr + 1), "p": (n_surface + 2) * (nr + 1)}
        if world == 1:
            assert len(self.node_x_u) == self.n_global_nodes["u"], "Q2 nodes of the shell were not identified one to one"
            assert len(self.node_x_p) == self.n_global_nodes["p"], "Q1 nodes of the shell were not identified one to one"
        self._part = {}
        self.n_u, self.n_p = 3 * len(self.node_x_u), len(self.node_x_p)
        cu = np.empty((self.n_cells, 81), dtype=_UINT)
        for c in range(3):
            cu[:, c * 27:(c + 1) * 27] = 3 * nodes_u + c
        self.cell_dofs_u = cu
        self.cell_dofs_p = nodes_p.astype(_UINT)
        self._geom = None

    # -- points ------------------------------------------------------------------------------------------------
    def points(self, ref, chunk=None, cells=None, tree=None):
        """Physical positions X of the reference points ref (npts, 3) of every local cell (or of the cells in `chunk`, or of
        the given cells of the given trees): (cells, npts, 3)."""
        if cells is None:
            sel = slice(None) if chunk is None else chunk
            cells, tree = self.cells[sel], self.tree[sel]
        ref = np.asarray(ref, dtype=np.float64)
        # the direction depends on (x, y) only: spherical average of the cell's 4 corner directions with Q1 weights
        xy, inv = np.unique(ref[:, :2], axis=0, return_inverse=True)
        inv = inv.ravel()
        L, i, j = self._lattice, cells[:, 0], cells[:, 1]
        C = np.stack([L[tree, j, i], L[tree, j, i + 1], L[tree, j + 1, i], L[tree, j + 1, i + 1]], axis=1)       # (cells, 4, 3)
        x, y = xy[:, 0], xy[:, 1]
        W = np.stack([(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y], axis=1)                            # (pts, 4)
        d = sph_avg(np.broadcast_to(C[:, None], (len(cells), len(xy), 4, 3)), np.broadcast_to(W[None], (len(cells), len(xy), 4)))
        r = self._ra[tree][:, None] + (cells[:, 2:3] + ref[None, :, 2]) / self.n * (self._rb - self._ra)[tree][:, None]
        return r[..., None] * d[:, inv]

    def _node_keys(self, cells, tree, ref):
        """Position keys (63-bit integers) of the reference points ref of the given cells: (cells, npts)."""
        scale = float(1 << (self.KEY_BITS - 2))
        out = np.empty((len(cells), len(ref)), dtype=np.int64)
        step = 1 << 14
        for s in range(0, len(cells), step):
            x = self.points(ref, cells=cells[s:s + step], tree=tree[s:s + step])
            q = np.rint((x + 1.0) * scale).astype(np.int64)
            out[s:s + step] = (q[..., 0] << (2 * self.KEY_BITS)) | (q[..., 1] << self.KEY_BITS) | q[..., 2]
        return out

    def _first_touch(self, family, npd_minus1, local_order, lex_order):
        """Number the nodes in first-touch order; nodes are identified by their quantised position."""
        scale = float(1 << (self.KEY_BITS - 2))

        def keys(order):
            return self._node_keys(self.cells, self.tree, order / float(npd_minus1))

        kfe = keys(local_order.astype(np.float64)).ravel()
        uniq, first = np.unique(kfe, return_index=True)
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
        del kfe
        klex = keys(lex_order.astype(np.float64))
        cell_nodes = rank[np.searchsorted(uniq, klex.ravel())].reshape(klex.shape)
        key_of_node = np.empty(len(uniq), dtype=np.int64)
        key_of_node[rank] = uniq
        self._keys[family] = key_of_node
        mask = (1 << self.KEY_BITS) - 1
        x = np.empty((len(uniq), 3))
        x[rank, 0] = (uniq >> (2 * self.KEY_BITS)) / scale - 1.0
        x[rank, 1] = ((uniq >> self.KEY_BITS) & mask) / scale - 1.0
        x[rank, 2] = (uniq & mask) / scale - 1.0
        return cell_nodes, x                                             # (positions are accurate to the key quantum, 2e-6)

    # -- constraints -------------------------------------------------------------------------------------------
    def _boundary_nodes(self, which):
        r = np.sqrt((self.node_x_u ** 2).sum(-1))
        on = np.zeros(len(r), bool)
        if which in ("bottom", "both"):
            on |= np.abs(r - self.R0) < 1e-4
        if which in ("top", "both"):
            on |= np.abs(r - self.R1) < 1e-4
        return on

    def constrained_u(self, bc):
        """Dirichlet DoFs.  'zero': all components on the inner and outer boundary (`Zero velocity boundary indicators = top,
        bottom`); 'zero bottom, tangential top': the inner boundary only (tests/shell_3d_gmg.prm) -- the outer boundary then
        carries the no-normal-flux lines of constraint_lines()."""
        if bc == "none":
            return np.zeros(0, dtype=_UINT)
        if bc not in ("zero", "zero bottom, tangential top", "tangential"):
            raise ValueError("shell boundaries: 'zero' | 'zero bottom, tangential top' | 'tangential' | 'none'")
        if bc == "tangential":                 # `Tangential velocity boundary indicators = top, bottom` (tests/gmg_no_normal_flux_only.prm)
            return np.zeros(0, dtype=_UINT)
        on = self._boundary_nodes("both" if bc == "zero" else "bottom")
        return np.flatnonzero(np.repeat(on[:, None], 3, axis=1).ravel()).astype(_UINT)

    def constraint_lines(self, family, bc):
        """No-normal-flux lines of `Tangential velocity boundary indicators = top` (or `top, bottom` for bc = 'tangential'; one rank): (dof, start, master, weight) in the
        format of smf_set_constraints.  As VectorTools::compute_no_normal_flux_constraints[_on_level] with
        use_manifold_for_normal = true (stokes_matrix_free_local_smoothing.cc:1736-1745): at every node of the outer boundary the
        normal is the radial direction, the component with the largest |n_k| (the first of them within 1e-9) is eliminated,
        u_k = -sum_{j != k} (n_j / n_k) u_j.  The node positions are evaluated exactly (node_x_u is accurate to the key quantum
        only)."""
        empty = (np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0))
        if family != "u" or bc not in ("zero bottom, tangential top", "tangential"):
            return empty
        assert self.world == 1, "no-normal-flux lines: one rank"
        top = self._boundary_nodes("top" if bc == "zero bottom, tangential top" else "both")
        cell_nodes = (self.cell_dofs_u[:, :27] // 3).astype(np.int64)
        cells = np.flatnonzero(top[cell_nodes].any(axis=1))                 # cells with a face on a no-normal-flux boundary
        ref = _lex(3, 3).astype(np.float64) / 2.0
        x = self.points(ref, cells=self.cells[cells], tree=self.tree[cells])   # (cells, 27, 3), exact
        nodes = cell_nodes[cells].ravel()
        uniq, first = np.unique(nodes, return_index=True)
        keep = top[uniq]
        uniq, xn = uniq[keep], x.reshape(-1, 3)[first[keep]]
        nrm = xn / np.sqrt((xn ** 2).sum(-1))[:, None]
        a = np.abs(nrm)
        k = np.argmax(a >= a.max(axis=1, keepdims=True) - 1e-9, axis=1)      # first component within 1e-9 of the largest
        others = np.array([[1, 2], [0, 2], [0, 1]])[k]
        rows = np.arange(len(uniq))
        dof = 3 * uniq + k
        master = (3 * uniq[:, None] + others).ravel()
        weight = (-nrm[rows[:, None], others] / nrm[rows, k][:, None]).ravel()
        start = 2 * np.arange(len(uniq) + 1)
        return dof.astype(np.int64), start.astype(np.int64), master.astype(np.int64), weight

    # -- geometry (MappingQ(degree) of the analytic tree map) --------------------------------------------------
    def geometry(self):
        """(xq (cells, 27, 3), jinv (cells, 27, 3, 3) with jinv[r][d] = d xi_r / d x_d, jxw (cells, 27))."""
        if self._geom is None:
            nodes, N, dN = mapping_tables(self.mapping_degree)
            ref = _lex(len(nodes), 3).astype(np.int64)
            ref = nodes[ref]                                             # support points of the mapping on the unit cell
            w = np.einsum("c,b,a->cba", GAUSS3_W, GAUSS3_W, GAUSS3_W).ravel()
            dN2 = np.ascontiguousarray(dN.transpose(0, 2, 1).reshape(81, -1))         # [(q, r)][k]
            xq = np.empty((self.n_cells, 27, 3)); jinv = np.empty((self.n_cells, 27, 3, 3)); jxw = np.empty((self.n_cells, 27))
            step = 1 << 13
            for s in range(0, self.n_cells, step):
                X = self.points(ref, slice(s, s + step))                # (cells, 125, 3)
                xq[s:s + step] = np.matmul(N[None], X)
                # J[d][r] = d x_d / d xihat_r (cell-local reference coordinates), as one batched GEMM
                J = np.matmul(dN2[None], X).reshape(len(X), 27, 3, 3).transpose(0, 1, 3, 2)
                jinv[s:s + step] = np.linalg.inv(J)
                jxw[s:s + step] = np.linalg.det(J) * w[None, :]
            assert (jxw > 0).all()
            self._geom = (xq, jinv, jxw)
        return self._geom

    def quadrature_points(self):
        return self.geometry()[0]

    def midpoint_geometry(self):
        """(position (cells, 3), det J (cells,)) at the midpoints of the local cells (QMidpoint of compute_pressure_scaling_factor)."""
        return self._midpoints_of(self.cells, self.tree)

    @property
    def jinv(self):
        return self.geometry()[1]

    @property
    def jxw(self):
        return self.geometry()[2]

    def dof_coords_u(self):
        return np.repeat(self.node_x_u, 3, axis=0)

    def surface_weights(self):
        """(integral of every Q1 pressure shape function over the outer boundary (n_p,), area of this rank's part of it):
        `Pressure normalization = surface` (helper_functions.cc:765-870; 3 x 3 Gauss points per boundary face)."""
        nodes, _, _ = mapping_tables(self.mapping_degree)
        V, D = _lagrange_tables(nodes, GAUSS3)
        V1, _ = _lagrange_tables(nodes, np.array([1.0]))
        ref = nodes[_lex(len(nodes), 3).astype(np.int64)]
        top = np.flatnonzero((self._rb[self.tree] == self.R1) & (self.cells[:, 2] == self.n - 1))
        w = np.zeros(self.n_p)
        area = 0.0
        nk = len(nodes)
        Tx = np.einsum("l,bj,ai->balji", V1[0], V, D).reshape(9, nk ** 3)          # d/dx of the face map at the 3 x 3 points
        Ty = np.einsum("l,bj,ai->balji", V1[0], D, V).reshape(9, nk ** 3)
        q1 = np.stack([1 - GAUSS3, GAUSS3], axis=1)
        phi = np.einsum("bj,ai->baji", q1, q1).reshape(9, 4)                         # face shape functions (x fastest)
        gw = np.einsum("b,a->ba", GAUSS3_W, GAUSS3_W).ravel()
        step = 1 << 13
        for s in range(0, len(top), step):
            sel = top[s:s + step]
            X = self.points(ref, cells=self.cells[sel], tree=self.tree[sel])
            da = np.sqrt((np.cross(np.matmul(Tx[None], X), np.matmul(Ty[None], X)) ** 2).sum(-1)) * gw[None, :]
            area += da.sum()
            np.add.at(w, self.cell_dofs_p[sel][:, 4:].astype(np.int64), da @ phi)
        return w, float(area)

    def global_midpoint_geometry(self):
        """midpoint_geometry() over ALL trees (the pressure scaling is a global reduction, helper_functions.cc:337-339)."""
        if self.world == 1:
            return self.midpoint_geometry()
        return ShellLevel._midpoints_of(self, np.tile(self._morton, (self.n_trees, 1)), np.repeat(np.arange(self.n_trees), len(self._morton)))

    def _midpoints_of(self, cells, tree):
        nodes, N, dN = mapping_tables(self.mapping_degree, np.array([0.5]))
        ref = nodes[_lex(len(nodes), 3).astype(np.int64)]
        xm = np.empty((len(cells), 3)); det = np.empty(len(cells))
        step = 1 << 14
        for s in range(0, len(cells), step):
            X = self.points(ref, cells=cells[s:s + step], tree=tree[s:s + step])
            xm[s:s + step] = np.einsum("k,ekd->ed", N[0], X)
            det[s:s + step] = np.linalg.det(np.einsum("kr,ekd->edr", dN[0], X))
        return xm, det

    # -- distributed layout ------------------------------------------------------------------------------------
    def _partition(self, family):
        """(global ids of the local nodes, {neighbour rank: local node indices shared with it, ordered by global id},
        number of ranks holding every local node).  The synthetic driver finds the sharers by generating the node keys of
        the other ranks' trees; ASPECT would read them from the Utilities::MPI::Partitioner of the level."""
        if family not in self._part:
            mine = self._keys[family]
            npd = 2 if family == "u" else 1
            ref = _lex(npd + 1, 3).astype(np.float64) / npd
            m = self._morton
            per_rank = []
            for q in range(self.world):
                if q == self.rank:
                    per_rank.append(np.unique(mine))
                else:
                    tr = np.repeat(np.arange(q * self.trees_per_rank, (q + 1) * self.trees_per_rank), len(m))
                    per_rank.append(np.unique(self._node_keys(np.tile(m, (self.trees_per_rank, 1)), tr, ref)))
            everyone = np.unique(np.concatenate(per_rank))
            assert len(everyone) == self.n_global_nodes[family], "nodes of the shell were not identified one to one across ranks"
            gid = np.searchsorted(everyone, mine)
            order = np.argsort(mine, kind="stable")
            lists, count = {}, np.ones(len(mine), dtype=np.int64)
            for q in range(self.world):
                if q == self.rank:
                    continue
                common = np.intersect1d(mine, per_rank[q], assume_unique=True)           # sorted by key = by global id
                if len(common):
                    idx = order[np.searchsorted(mine[order], common)]
                    lists[q] = idx.astype(_UINT)
                    count[idx] += 1
            self._part[family] = (gid, lists, count)
        return self._part[family]

    def global_node_ids(self, family):
        """Partition-independent id of every local node (rank of its position key among all nodes of the level)."""
        return self._partition(family)[0]

    def exchange_lists(self, family):
        """{neighbour rank: local node indices shared with it, ordered by global node id (identical on both sides)}."""
        return self._partition(family)[1]

    def sharer_counts(self, family):
        """Number of ranks that hold each local node (weights of the distributed inner products are 1 / this)."""
        return self._partition(family)[2]


class ShellHierarchy:
    """Levels 0..R of the uniformly refined shell; active mesh = level R."""

    def __init__(self, refinement, inner_radius=0.54, outer_radius=1.0, mapping_degree=4, rank=0, world=1, n_coarse=96):
        self.dim, self.R = 3, refinement
        self.rank, self.world = rank, world
        self.levels = [ShellLevel(l, inner_radius, outer_radius, mapping_degree, rank, world, n_coarse) for l in range(refinement + 1)]
        self.active = self.levels[-1]

    def children(self, level):
        npar = self.levels[level - 1].n_cells
        return (np.arange(npar, dtype=np.int64)[:, None] * 8 + np.arange(8)[None, :]).astype(np.int32)
