"""Adaptively refined box meshes with hanging nodes (oracle; test infrastructure) -- groundwork for SURVEY 8(a) row a12.

What the reference gets from deal.II on an adaptively refined `parallel::distributed::Triangulation` of the box
(BASELINE cfg 3 / cfg 4 refine adaptively; /root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:
1487-1633 builds the constraints, :1695-1803 the level constraints):

* mesh: one tree (or `repetitions` trees of a subdivided box, tree by tree), leaves = active cells, 2:1 balanced across faces, edges and corners (p4est `P4EST_CONNECT_FULL`), active
  cells in depth-first (z-) order;
* active DoFs: `distribute_dofs` + `DoFRenumbering::hierarchical` = first touch over the active cells in that order, FE-local
  order inside a cell; hanging nodes carry DoFs like any other node;
* hanging-node constraints (`DoFTools::make_hanging_node_constraints`): a node of a fine cell that lies on a face / edge
  of a coarser neighbour without being one of its nodes equals the coarse cell's interpolant there -- Q2 in 1D:
  u(1/4) = 3/8 u0 + 3/4 u1 - 1/8 u2, u(3/4) = -1/8 u0 + 3/4 u1 + 3/8 u2; Q1: u(1/2) = (u0 + u1) / 2;
* operator convention of `MatrixFreeOperators::Base` (App. A.1): `read_dof_values` resolves the constraints (hanging nodes
  read the weighted master DoFs, Dirichlet DoFs read 0), `distribute_local_to_global` applies their transpose, constrained
  rows get a unit diagonal: A_mf = C^T A C + I_constrained with the constraint matrix C.

AdaptiveBox restates the ACTIVE level (mesh, numbering, constraints, constrained operators); AdaptiveHierarchy /
AdaptiveVCycle restate local smoothing on such a mesh (level meshes, refinement-edge indices, interface operators as
edge_out / edge_in of the V-cycle, copy_to_mg / copy_from_mg), block_gmg_solve the cheap phase of the solve on top.
Checks (tests/test_adaptive_oracle.py): without adaptive refinement everything reduces to the uniform restatement (same
index maps, same matrices, the reference's 21 iterations of tests/nsinker_gmg_gc); Q2 / Q1 polynomials satisfy the hanging
constraints exactly; rigid motions are in the null space; the GMG-preconditioned solve on an adaptive mesh converges to the
direct solution of the constrained system, with the iteration count of the uniform mesh for a constant viscosity.
Pinned by the reference golden tests/muparser_temperature_example (oracle/muparser_example.py: meshes from a minimum refinement
function alone; cells, levels, pressure DoFs exact, 6 / 9 / 9 iterations reproduced as 7 / 9 / 10); the other adaptive goldens of the
reference depend on Kelly refinement indicators, which are outside this path.
"""
import itertools

import numpy as np
import scipy.sparse as sp

from . import fe


def _constrained_components(key, hi, bc, dim):
    """Velocity components that carry a Dirichlet value at the lattice point `key`: bc = 'zero' | 'tangential' | 'none' for all
    faces alike, or a dict {(direction, side): 'zero' | 'tangential' | 'none'} per face (side 0: low, 1: high)."""
    out = set()
    for d in range(dim):
        for side, v in ((0, 0), (1, hi[d])):
            if key[d] != v:
                continue
            kind = bc.get((d, side), "none") if isinstance(bc, dict) else bc
            if kind == "zero":
                out.update(range(dim))
            elif kind == "tangential":
                out.add(d)
    return out


def _children(cell, dim):
    l, ijk = cell
    return [(l + 1, tuple(2 * ijk[d] + ((k >> d) & 1) for d in range(dim))) for k in range(1 << dim)]


class AdaptiveBox:
    """Box of `repetitions` coarse cells (trees; default one) of edge length `coarse_size` (default 1: the unit box);
    `leaves` = set of active cells (level, (i, j[, k])) with 0 <= i < repetitions[0] * 2^level etc."""

    def __init__(self, dim, leaves, repetitions=None, coarse_size=1.0):
        self.dim = dim
        self.reps = tuple(repetitions) if repetitions is not None else (1,) * dim
        self.coarse_size = np.full(dim, float(coarse_size)) if np.ndim(coarse_size) == 0 else np.asarray(coarse_size, float)
        self.leaves = set(leaves)
        self.Lmax = max(l for l, _ in self.leaves)
        self._balance_check()
        # p4est order: tree by tree (trees of GridGenerator::subdivided_hyper_rectangle: x fastest), inside a tree depth-first
        # = Morton order of the anchors at the finest resolution
        def morton(cell):
            l, ijk = cell
            s = self.Lmax - l
            tree = [ijk[d] >> l for d in range(dim)]
            t = sum(tree[d] * int(np.prod(self.reps[:d])) for d in range(dim))
            m = 0
            for b in range(self.Lmax):
                for d in range(dim):
                    m |= ((((ijk[d] - (tree[d] << l)) << s) >> b) & 1) << (dim * b + d)
            return (t, m)
        self.cells = sorted(self.leaves, key=morton)
        self.n_cells = len(self.cells)
        self._number()
        self._hanging()

    # -- construction helpers ---------------------------------------------------------------------------------
    @classmethod
    def refined(cls, dim, base_level, marker, n_adaptive, repetitions=None, coarse_size=1.0):
        """Uniform refinement to `base_level`, then `n_adaptive` rounds: refine every active cell with marker(center, level)
        true and whatever the 2:1 balance (faces, edges, corners) then requires."""
        reps = tuple(repetitions) if repetitions is not None else (1,) * dim
        leaves = {(base_level, ijk) for ijk in itertools.product(*[range(reps[d] << base_level) for d in range(dim)])}
        for _ in range(n_adaptive):
            size = np.full(dim, float(coarse_size)) if np.ndim(coarse_size) == 0 else np.asarray(coarse_size, float)
            marked = {c for c in leaves if marker((np.array(c[1]) + 0.5) * size / (1 << c[0]), c[0])}
            leaves = cls._refine(dim, leaves, marked, reps)
        return cls(dim, leaves, reps, coarse_size)

    @staticmethod
    def _neighbours(cell, dim, reps=None):
        l, ijk = cell
        reps = reps if reps is not None else (1,) * dim
        out = []
        for off in itertools.product((-1, 0, 1), repeat=dim):
            if any(off):
                q = tuple(ijk[d] + off[d] for d in range(dim))
                if all(0 <= q[d] < (reps[d] << l) for d in range(dim)):
                    out.append((l, q))
        return out

    @classmethod
    def _refine(cls, dim, leaves, marked, reps=None):
        leaves = set(leaves)
        queue = list(marked)
        while queue:
            c = queue.pop()
            if c not in leaves:
                continue
            leaves.remove(c)
            kids = _children(c, dim)
            leaves.update(kids)
            # every same-level neighbour position of c must be covered by a leaf of level >= l: coarser ones are refined
            for nb in cls._neighbours(c, dim, reps):
                l, ijk = nb
                if l == 0:
                    continue
                anc = (l - 1, tuple(i // 2 for i in ijk))
                if anc in leaves:
                    queue.append(anc)
        return leaves

    def _covering_leaf(self, level, ijk):
        """The active cell that contains the level-`level` cell position ijk (itself, or an ancestor), else None (refined further)."""
        l, q = level, tuple(ijk)
        while l >= 0:
            if (l, q) in self.leaves:
                return (l, q)
            l, q = l - 1, tuple(i // 2 for i in q)
        return None

    def _balance_check(self):
        for c in self.leaves:
            for nb in self._neighbours(c, self.dim, self.reps):
                cov = self._covering_leaf(*nb)
                assert cov is None or cov[0] >= c[0] - 1, f"mesh is not 2:1 balanced at {c}"

    # -- geometry -----------------------------------------------------------------------------------------------
    def node_key(self, cell, loc, npd):
        """Integer position of local node loc (entries 0..npd-1) of `cell` on the finest lattice of that element family."""
        l, ijk = cell
        s = 1 << (self.Lmax - l)
        return tuple(((npd - 1) * ijk[d] + loc[d]) * s for d in range(self.dim))

    def cell_size(self, cell):
        """edge lengths (dim,) of a cell"""
        return self.coarse_size / (1 << cell[0])

    def quadrature_points(self):
        t = fe.tensor_tables(self.dim)
        return np.array([[(np.array(ijk) + xq) * self.coarse_size / (1 << l) for xq in t["xq"]] for l, ijk in self.cells])

    # -- numbering ----------------------------------------------------------------------------------------------
    def _number(self):
        dim = self.dim
        ord2, ord1 = fe.fe_q_local_order(2, dim), fe.fe_q_local_order(1, dim)
        lex2, lex1 = fe.lex_nodes(3, dim), fe.lex_nodes(2, dim)
        nu, nprs = {}, {}
        cu = np.zeros((self.n_cells, dim * 3 ** dim), dtype=np.int64)
        cp = np.zeros((self.n_cells, 2 ** dim), dtype=np.int64)
        for m, cell in enumerate(self.cells):
            for loc in ord2:
                nu.setdefault(self.node_key(cell, loc, 3), len(nu))
            for a, loc in enumerate(lex2):
                k = nu[self.node_key(cell, loc, 3)]
                for c in range(dim):
                    cu[m, c * 3 ** dim + a] = dim * k + c
            for loc in ord1:
                nprs.setdefault(self.node_key(cell, loc, 2), len(nprs))
            for a, loc in enumerate(lex1):
                cp[m, a] = nprs[self.node_key(cell, loc, 2)]
        self.cell_dofs_u, self.cell_dofs_p = cu, cp
        self.node_u, self.node_p = nu, nprs
        self.n_u, self.n_p = dim * len(nu), len(nprs)

    # -- hanging nodes ------------------------------------------------------------------------------------------
    def _hanging_of(self, npd, nodes, shape):
        """{node index: [(master node index, weight)]} for the scalar element with npd nodes per direction."""
        dim, per = self.dim, npd - 1
        by_cell = {c: {self.node_key(c, loc, npd) for loc in fe.lex_nodes(npd, dim)} for c in self.cells}
        lex = fe.lex_nodes(npd, dim)
        cons = {}
        for cell in self.cells:
            l = cell[0]
            if l == 0:
                continue
            span = per << (self.Lmax - (l - 1))               # lattice length of a cell one level coarser
            for loc in lex:
                key = self.node_key(cell, loc, npd)
                if nodes[key] in cons:
                    continue
                # coarser active cells that contain the point
                cands = itertools.product(*[{max(0, min(key[d] // span, (self.reps[d] << (l - 1)) - 1)),
                                             max(0, min((key[d] - 1) // span, (self.reps[d] << (l - 1)) - 1))}
                                            for d in range(dim)])
                for q in cands:
                    coarse = (l - 1, tuple(q))
                    if coarse not in self.leaves or key in by_cell[coarse]:
                        continue
                    x = [(key[d] - q[d] * span) / span for d in range(dim)]
                    if not all(0.0 <= v <= 1.0 for v in x):
                        continue
                    w = []
                    for cl in lex:
                        v = np.prod([shape(x[d])[cl[d]] for d in range(dim)])
                        if abs(v) > 1e-14:
                            w.append((nodes[self.node_key(coarse, cl, npd)], float(v)))
                    cons[nodes[key]] = w
                    break
        # resolve chains (a master that is itself constrained), as AffineConstraints::close() does
        changed = True
        while changed:
            changed = False
            for k, w in list(cons.items()):
                if any(m in cons for m, _ in w):
                    acc = {}
                    for m, v in w:
                        for m2, v2 in (cons[m] if m in cons else [(m, 1.0)]):
                            acc[m2] = acc.get(m2, 0.0) + v * v2
                    cons[k] = sorted(acc.items())
                    changed = True
        return cons

    def _hanging(self):
        q2 = lambda x: [(1 - x) * (1 - 2 * x), 4 * x * (1 - x), x * (2 * x - 1)]
        q1 = lambda x: [1 - x, x]
        self.hanging_u = self._hanging_of(3, self.node_u, q2)       # per velocity NODE (all components alike)
        self.hanging_p = self._hanging_of(2, self.node_p, q1)

    # -- constraint matrices ------------------------------------------------------------------------------------
    def dirichlet_u(self, bc):
        """Constrained velocity DoFs on the box boundary: 'zero' | 'tangential' | 'none' (as oracle/mesh.py), or per face
        (_constrained_components)."""
        if bc == "none":
            return np.zeros(0, dtype=np.int64)
        hi = [self.reps[d] * (2 << self.Lmax) for d in range(self.dim)]
        out = []
        for key, k in self.node_u.items():
            out += [self.dim * k + c for c in _constrained_components(key, hi, bc, self.dim)]
        return np.array(sorted(out), dtype=np.int64)

    def constraint_matrices(self, bc):
        """(C_u, C_p, constrained velocity DoFs, constrained pressure DoFs): x_full = C x with the master DoFs as columns
        (constrained columns are empty); Dirichlet DoFs read zero, hanging DoFs read their masters."""
        dim = self.dim
        dirichlet = set(self.dirichlet_u(bc).tolist())
        rows, cols, vals = [], [], []
        cons_u = set(dirichlet)
        for k in range(self.n_u // dim):
            for c in range(dim):
                i = dim * k + c
                if i in dirichlet:
                    continue
                if k in self.hanging_u:
                    cons_u.add(i)
                    for m, w in self.hanging_u[k]:
                        if dim * m + c not in dirichlet:
                            rows.append(i); cols.append(dim * m + c); vals.append(w)
                else:
                    rows.append(i); cols.append(i); vals.append(1.0)
        Cu = sp.coo_matrix((vals, (rows, cols)), shape=(self.n_u, self.n_u)).tocsr()
        rows, cols, vals = [], [], []
        for k in range(self.n_p):
            if k in self.hanging_p:
                for m, w in self.hanging_p[k]:
                    rows.append(k); cols.append(m); vals.append(w)
            else:
                rows.append(k); cols.append(k); vals.append(1.0)
        Cp = sp.coo_matrix((vals, (rows, cols)), shape=(self.n_p, self.n_p)).tocsr()
        return Cu, Cp, np.array(sorted(cons_u), dtype=np.int64), np.array(sorted(self.hanging_p), dtype=np.int64)

    # -- operators ----------------------------------------------------------------------------------------------
    def assemble(self, eta, pressure_scaling=1.0, bc="zero"):
        """Constrained active-level operators in the MatrixFreeOperators::Base convention:
        returns dict(A, BT, M, K = [A BT; BT^T 0]) with unit diagonals on the constrained rows of A and M.
        eta: (n_cells,) or (n_cells, n_q)."""
        dim = self.dim
        t = fe.tensor_tables(dim)
        nn, nq = 3 ** dim, len(t["w"])
        eta = np.asarray(eta, float)
        eta = np.repeat(eta[:, None], nq, axis=1) if eta.ndim == 1 else eta
        Ra, Ca, Va, Rb, Cb, Vb, Rm, Cm, Vm = ([] for _ in range(9))
        for m, cell in enumerate(self.cells):
            h = self.cell_size(cell)
            vol = float(np.prod(h))
            G = np.zeros((nq, dim * nn, dim, dim))
            for c in range(dim):
                G[:, c * nn:(c + 1) * nn, c, :] = t["dphi2"] / h[None, None, :]
            E = 0.5 * (G + np.swapaxes(G, 2, 3))
            coef = eta[m] * t["w"] * vol
            Ke = 2.0 * np.einsum("q,qiab,qjab->ij", coef, E, E)
            du = self.cell_dofs_u[m]
            Ra.append(np.repeat(du, len(du))); Ca.append(np.tile(du, len(du))); Va.append(Ke.ravel())
            div = np.trace(G, axis1=2, axis2=3)
            Be = -pressure_scaling * np.einsum("q,qi,qj->ij", t["w"] * vol, div, t["phi1"])
            dp = self.cell_dofs_p[m]
            Rb.append(np.repeat(du, len(dp))); Cb.append(np.tile(dp, len(du))); Vb.append(Be.ravel())
            Me = np.einsum("q,qi,qj->ij", pressure_scaling ** 2 / eta[m] * t["w"] * vol, t["phi1"], t["phi1"])
            Rm.append(np.repeat(dp, len(dp))); Cm.append(np.tile(dp, len(dp))); Vm.append(Me.ravel())
        cat = np.concatenate
        A = sp.coo_matrix((cat(Va), (cat(Ra), cat(Ca))), shape=(self.n_u, self.n_u)).tocsr()
        BT = sp.coo_matrix((cat(Vb), (cat(Rb), cat(Cb))), shape=(self.n_u, self.n_p)).tocsr()
        M = sp.coo_matrix((cat(Vm), (cat(Rm), cat(Cm))), shape=(self.n_p, self.n_p)).tocsr()
        Cu, Cp, cons_u, cons_p = self.constraint_matrices(bc)
        Iu = sp.coo_matrix((np.ones(len(cons_u)), (cons_u, cons_u)), shape=A.shape).tocsr()
        Ip = sp.coo_matrix((np.ones(len(cons_p)), (cons_p, cons_p)), shape=M.shape).tocsr()
        Ac = (Cu.T @ A @ Cu + Iu).tocsr()
        BTc = (Cu.T @ BT @ Cp).tocsr()
        Mc = (Cp.T @ M @ Cp + Ip).tocsr()
        # the Stokes operator has a unit diagonal on the constrained pressure rows, too (MatrixFreeOperators::Base)
        K = sp.bmat([[Ac, BTc], [BTc.T, Ip]], format="csr")
        return dict(A=Ac, BT=BTc, M=Mc, K=K, Cu=Cu, Cp=Cp, constrained_u=cons_u, constrained_p=cons_p,
                    A_raw=A, BT_raw=BT, M_raw=M)


# ----------------------------------------------------------------------------------------------------------------------
# Local smoothing on the adaptive mesh: level meshes, refinement-edge indices, interface operators, copy indices
# ----------------------------------------------------------------------------------------------------------------------
class AdaptiveHierarchy:
    """Level hierarchy of local smoothing on an AdaptiveBox, as deal.II builds it for
    `Stokes GMG type = local smoothing` (stokes_matrix_free_local_smoothing.cc:1559-1850):

    * level l = ALL cells of level l (active or refined further), `distribute_mg_dofs` numbering (first touch, z-order);
    * MGConstrainedDoFs: boundary indices (Dirichlet) and refinement-edge indices = level DoFs on faces between level-l
      cells and coarser active cells;
    * the level MatrixFree objects carry only the boundary constraints (:1710-1718); the refinement edge is handled by
      MatrixFreeOperators::Base: with A_bc the level operator under the boundary constraints and E the
      edge mask,   A_l = (I - E) A_bc (I - E) + E,   vmult_interface_down = E A_bc (I - E),   vmult_interface_up = (I - E) A_bc E;
    * MGTransfer copy indices: (active DoF, level DoF) for the DoFs of the active cells of level l that are not on the
      level's refinement edge (those are copied on the coarser level, where they are ordinary DoFs);
    * viscosity on the levels: interpolate_to_mg of the DGQ0 vector = the cell's own value where it is active, else the mean
      of its children (:220-226).
    """

    def __init__(self, box, bc="zero"):
        self.box, self.dim, self.bc = box, box.dim, bc
        dim = self.dim
        self.n_levels = box.Lmax + 1
        # level cells
        cells = [set() for _ in range(self.n_levels)]
        for l, ijk in box.leaves:
            q = ijk
            for ll in range(l, -1, -1):
                cells[ll].add((ll, q))
                q = tuple(i // 2 for i in q)

        reps = box.reps

        def morton(cell):
            l, ijk = cell
            tree = [ijk[d] >> l for d in range(dim)]
            t = sum(tree[d] * int(np.prod(reps[:d])) for d in range(dim))
            m = 0
            for b in range(l):
                for d in range(dim):
                    m |= (((ijk[d] - (tree[d] << l)) >> b) & 1) << (dim * b + d)
            return (t, m)
        self.cells = [sorted(c, key=morton) for c in cells]
        self.index = [{c: k for k, c in enumerate(cl)} for cl in self.cells]
        self.levels = [self._number_level(l) for l in range(self.n_levels)]
        self._copy_indices()

    def _key(self, l, cell, loc, npd):
        s = 1 << (self.box.Lmax - l)
        return tuple(((npd - 1) * cell[1][d] + loc[d]) * s for d in range(self.dim))

    def _number_level(self, l):
        dim, box = self.dim, self.box
        ord2, ord1 = fe.fe_q_local_order(2, dim), fe.fe_q_local_order(1, dim)
        lex2, lex1 = fe.lex_nodes(3, dim), fe.lex_nodes(2, dim)
        nu, npn = {}, {}
        cl = self.cells[l]
        cu = np.zeros((len(cl), dim * 3 ** dim), dtype=np.int64)
        cp = np.zeros((len(cl), 2 ** dim), dtype=np.int64)
        for m, cell in enumerate(cl):
            for loc in ord2:
                nu.setdefault(self._key(l, cell, loc, 3), len(nu))
            for a, loc in enumerate(lex2):
                k = nu[self._key(l, cell, loc, 3)]
                for c in range(dim):
                    cu[m, c * 3 ** dim + a] = dim * k + c
            for loc in ord1:
                npn.setdefault(self._key(l, cell, loc, 2), len(npn))
            for a, loc in enumerate(lex1):
                cp[m, a] = npn[self._key(l, cell, loc, 2)]

        def edge_nodes(nodes, per):
            """nodes of the level that touch a level-l cell position inside the box that is not part of the level mesh"""
            span = per << (box.Lmax - l)
            n = [box.reps[d] << l for d in range(dim)]
            out = set()
            for key, k in nodes.items():
                touching = itertools.product(*[{max(0, min(key[d] // span, n[d] - 1)), max(0, min((key[d] - 1) // span, n[d] - 1))}
                                               for d in range(dim)])
                if any((l, tuple(q)) not in self.index[l] for q in touching):
                    out.add(k)
            return out
        eu, ep = edge_nodes(nu, 2), edge_nodes(npn, 1)
        hi = [box.reps[d] * (2 << box.Lmax) for d in range(dim)]
        bnd = []
        if self.bc != "none":
            for key, k in nu.items():
                bnd += [dim * k + c for c in _constrained_components(key, hi, self.bc, dim)]
        return dict(cell_dofs_u=cu, cell_dofs_p=cp, node_u=nu, node_p=npn, n_u=dim * len(nu), n_p=len(npn), h=box.coarse_size / (1 << l),
                    boundary_u=np.array(sorted(bnd), dtype=np.int64),
                    edge_u=np.array(sorted(dim * k + c for k in eu for c in range(dim)), dtype=np.int64),
                    edge_p=np.array(sorted(ep), dtype=np.int64))

    def _copy_indices(self):
        box, dim = self.box, self.dim
        lex2, lex1 = fe.lex_nodes(3, dim), fe.lex_nodes(2, dim)
        self.copy_u, self.copy_p = [], []
        for l, L in enumerate(self.levels):
            eu, ep = set(L["edge_u"].tolist()), set(L["edge_p"].tolist())
            pu, pp = {}, {}
            for cell in self.cells[l]:
                if cell not in box.leaves:
                    continue
                for loc in lex2:
                    ka, kl = box.node_u[box.node_key(cell, loc, 3)], L["node_u"][self._key(l, cell, loc, 3)]
                    for c in range(dim):
                        if dim * kl + c not in eu:
                            pu[dim * ka + c] = dim * kl + c
                for loc in lex1:
                    ka, kl = box.node_p[box.node_key(cell, loc, 2)], L["node_p"][self._key(l, cell, loc, 2)]
                    if kl not in ep:
                        pp[ka] = kl
            self.copy_u.append((np.array(sorted(pu), dtype=np.int64), np.array([pu[k] for k in sorted(pu)], dtype=np.int64)))
            self.copy_p.append((np.array(sorted(pp), dtype=np.int64), np.array([pp[k] for k in sorted(pp)], dtype=np.int64)))

    # -- coefficient tables ----------------------------------------------------------------------------------------
    def level_viscosity(self, eta_active):
        """DGQ0 value of every level cell from the active cells' values (n_cells,)."""
        val = {c: float(eta_active[k]) for k, c in enumerate(self.box.cells)}
        out = [None] * self.n_levels
        for l in range(self.n_levels - 1, -1, -1):
            v = np.zeros(len(self.cells[l]))
            for k, c in enumerate(self.cells[l]):
                if c not in val:
                    val[c] = float(np.mean([val[ch] for ch in _children(c, self.dim)]))
                v[k] = val[c]
            out[l] = v
        return out

    # -- level operators ----------------------------------------------------------------------------------------
    def level_operators(self, eta_levels, pressure_scaling):
        """per level: dict(A, M) smoother operators, (A_down, A_up, M_down, M_up) interface operators."""
        dim = self.dim
        t = fe.tensor_tables(dim)
        nn = 3 ** dim
        ops = []
        for l, L in enumerate(self.levels):
            h = np.asarray(L["h"], float)
            vol = float(np.prod(h))
            G = np.zeros((len(t["w"]), dim * nn, dim, dim))
            for c in range(dim):
                G[:, c * nn:(c + 1) * nn, c, :] = t["dphi2"] / h[None, None, :]
            E = 0.5 * (G + np.swapaxes(G, 2, 3))
            K1 = 2.0 * np.einsum("q,qiab,qjab->ij", t["w"] * vol, E, E)
            M1 = np.einsum("q,qi,qj->ij", t["w"] * vol, t["phi1"], t["phi1"])
            cu, cp = L["cell_dofs_u"], L["cell_dofs_p"]
            eta = eta_levels[l]
            A = sp.coo_matrix(((eta[:, None, None] * K1[None]).ravel(),
                               (np.repeat(cu, cu.shape[1], axis=1).ravel(), np.tile(cu, (1, cu.shape[1])).ravel())),
                              shape=(L["n_u"], L["n_u"])).tocsr()
            M = sp.coo_matrix((((pressure_scaling ** 2 / eta)[:, None, None] * M1[None]).ravel(),
                               (np.repeat(cp, cp.shape[1], axis=1).ravel(), np.tile(cp, (1, cp.shape[1])).ravel())),
                              shape=(L["n_p"], L["n_p"])).tocsr()
            # boundary constraints of the level MatrixFree: reads 0, no writes, unit diagonal
            keep = np.ones(L["n_u"]); keep[L["boundary_u"]] = 0.0
            Db = sp.diags(keep)
            A_bc = (Db @ A @ Db + sp.diags(1.0 - keep)).tocsr()

            def split(Mat, edge, n):
                e = np.zeros(n); e[edge] = 1.0
                Ed, Id = sp.diags(e), sp.diags(1.0 - e)
                return (Id @ Mat @ Id + Ed).tocsr(), (Ed @ Mat @ Id).tocsr(), (Id @ Mat @ Ed).tocsr()
            A_l, A_down, A_up = split(A_bc, L["edge_u"], L["n_u"])
            M_l, M_down, M_up = split(M, L["edge_p"], L["n_p"])
            ops.append(dict(A=A_l, A_down=A_down, A_up=A_up, M=M_l, M_down=M_down, M_up=M_up))
        return ops

    def prolongation(self, l, space):
        """P (n_l x n_{l-1}): per refined parent the tensor-product child-patch interpolation, 1/valence weights over the
        patches that exist; coarse Dirichlet DoFs are zeroed first (velocity)."""
        dim = self.dim
        F, C = self.levels[l], self.levels[l - 1]
        if space == "u":
            P1, npd, ncomp, cf, cc, nf, nc = fe.P2_1D, 3, dim, F["cell_dofs_u"], C["cell_dofs_u"], F["n_u"], C["n_u"]
        else:
            P1, npd, ncomp, cf, cc, nf, nc = fe.P1_1D, 2, 1, F["cell_dofs_p"], C["cell_dofs_p"], F["n_p"], C["n_p"]
        npatch = 2 * npd - 1
        Ploc = P1
        for _ in range(dim - 1):
            Ploc = np.kron(P1, Ploc)
        child_nodes = fe.lex_nodes(npd, dim)
        nn = npd ** dim
        rows, cols, vals = [], [], []
        for pk, par in enumerate(self.cells[l - 1]):
            kids = _children(par, dim)
            if kids[0] not in self.index[l]:
                continue
            for c in range(ncomp):
                patch = np.zeros(npatch ** dim, dtype=np.int64)
                for k, kid in enumerate(kids):
                    off = [(k >> d) & 1 for d in range(dim)]
                    fk = self.index[l][kid]
                    for a, loc in enumerate(child_nodes):
                        pi = sum((loc[d] + (npd - 1) * off[d]) * npatch ** d for d in range(dim))
                        patch[pi] = cf[fk, c * nn + a]
                cd = cc[pk, c * nn:(c + 1) * nn]
                rows.append(np.repeat(patch, nn)); cols.append(np.tile(cd, npatch ** dim)); vals.append(Ploc.ravel())
        rows, cols, vals = map(np.concatenate, (rows, cols, vals))
        valence = np.bincount(rows, minlength=nf).astype(float) / nn
        vals = vals / valence[rows]
        P = sp.coo_matrix((vals, (rows, cols)), shape=(nf, nc)).tocsr()
        if space == "u" and len(C["boundary_u"]):
            keep = np.ones(nc); keep[C["boundary_u"]] = 0.0
            P = P @ sp.diags(keep)
        return P


class AdaptiveVCycle:
    """PreconditionMG::vmult on the adaptive hierarchy: copy_to_mg, Multigrid::level_v_step with edge matrices
    (set_edge_matrices(mg_interface, mg_interface), local_smoothing.cc:903,911), copy_from_mg."""

    def __init__(self, hier, ops, which, steps):
        from .solver import Chebyshev
        self.h, self.steps = hier, steps
        key = "A" if which == "u" else "M"
        self.Aop = [o[key] for o in ops]
        self.down = [o[key + "_down"] for o in ops]
        self.up = [o[key + "_up"] for o in ops]
        self.P = [None] + [hier.prolongation(l, which) for l in range(1, hier.n_levels)]
        self.copy = hier.copy_u if which == "u" else hier.copy_p
        self.n = [L["n_u"] if which == "u" else L["n_p"] for L in hier.levels]
        self.smoothers = []
        for l in range(hier.n_levels):
            A = self.Aop[l]
            s = Chebyshev(lambda x, A=A: A @ x, 1.0 / A.diagonal(), 4 if l > 0 else 8, 15.0 if l > 0 else 1e-3, 10 if l > 0 else 100)
            s.estimate_eigenvalues()
            self.smoothers.append(s)

    def vmult(self, src):
        L = self.h.n_levels - 1
        defect = [np.zeros(n) for n in self.n]
        for l, (ia, il) in enumerate(self.copy):            # copy_to_mg
            defect[l][il] = src[ia]
        sol = [None] * (L + 1)
        self._sol = sol                                     # (copy_from_mg needs the solution of every level)
        self._run(L, defect)
        dst = np.zeros_like(src)
        for l, (ia, il) in enumerate(self.copy):            # copy_from_mg
            dst[ia] = sol[l][il]
        return dst

    def _run(self, l, defect):
        """Multigrid::level_v_step"""
        sm = self.smoothers[l]
        x = sm.vmult(defect[l])
        for _ in range(1, self.steps):
            x = sm.step(x, defect[l])
        if l > 0:
            t = defect[l] - (self.Aop[l] @ x + self.down[l] @ x)      # matrix->vmult, edge_out->vmult_add
            defect[l - 1] = defect[l - 1] + self.P[l].T @ t           # restrict_and_add onto the copy_to_mg part
            xc = self._run(l - 1, defect)
            x = x + self.P[l] @ xc
            d = defect[l] - self.up[l] @ x                            # defect -= edge_in.Tvmult(solution)
            for _ in range(self.steps):
                x = sm.step(x, d)
        self._sol[l] = x
        return x


def block_gmg_solve(box, eta_active, rhs, bc="zero", pressure_scaling=1.0, tol_rel=1e-6, restart=100, max_its=1000):
    """Cheap phase of StokesMatrixFreeHandler::solve() on an adaptive mesh: right-preconditioned GMRES on the constrained
    active operator with the block preconditioner built from the two adaptive V-cycles."""
    from .solver import gmres_right
    ops = box.assemble(eta_active, pressure_scaling, bc)
    hier = AdaptiveHierarchy(box, bc)
    lops = hier.level_operators(hier.level_viscosity(eta_active), pressure_scaling)
    mgA, mgS = AdaptiveVCycle(hier, lops, "u", 1), AdaptiveVCycle(hier, lops, "p", 4)
    nu = box.n_u
    K, BT = ops["K"], ops["BT"]

    def prec(r):
        zp = -mgS.vmult(r[nu:])
        zu = mgA.vmult(r[:nu] - BT @ zp)
        return np.concatenate([zu, zp])
    b = rhs.copy()
    b[ops["constrained_u"]] = 0.0
    b[nu + ops["constrained_p"]] = 0.0
    tol = tol_rel * np.linalg.norm(b)
    x, its, res, ok = gmres_right(lambda v: K @ v, prec, b, np.zeros_like(b), tol, restart, max_its)
    return dict(x=x, iterations=its, residual=res, converged=ok, ops=ops, hierarchy=hier)
