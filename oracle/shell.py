"""Spherical-shell mesh hierarchy with curved cells (oracle; test infrastructure).

Literal, loop-based restatement of the mesh the synthetic shell driver documents (BASELINE config 3; the reference
builds it with `GridGenerator::hyper_shell(..., 96, true)`, /root/reference/source/geometry_model/spherical_shell.cc:361-373,
and maps it with MappingQ(4), source/simulator/core.cc:118-119):

* coarse mesh: the 12 rhombic faces of the rhombic dodecahedron, one per cube edge (edge direction k = 0, 1, 2; inside k
  the signs (s_j, s_i) of the two other coordinates j = k+2, i = k+1 run (-,-), (-,+), (+,-), (+,+)); face corners
  v00 = s_i e_i, v11 = s_j e_j and the two cube corners, ordered so that (xi, eta, outward) is right-handed; every face is
  split 2 x 2 at its projected edge midpoints / centre; 2 radial layers [R0, Rm], [Rm, R1]; coarse cell order:
  face, layer, sub-quad y, sub-quad x  ->  96 cells, 150 vertices;
* tree map X(xi, eta, zeta) = (r_a + zeta (r_b - r_a)) normalize(bilinear blend of the corner directions);
* level-l cells: tree by tree, Morton order inside the tree; DoFs: first touch, FE-local order;
* geometry: x(xihat) = sum_k X_k L_k(xihat) with the 5^3 tensor Gauss-Lobatto support points of the cell,
  J = dx/dxihat, tables jinv[cell][q][r][d] = d xihat_r / d x_d and jxw[cell][q] = det J w_q at the 3^3 Gauss points.

parity unpinned at the geometry level: deal.II is not available, so the coarse-cell order and the interior point placement
cannot be compared with the reference; the topology is pinned by the reference golden tests/spherical_nsinker_gmg_gc
(20,790 + 970 DoFs at refinement 1, coarse sizes 2910 / 150).
"""
import math

import numpy as np

from . import fe

LOBATTO5 = [0.0, 0.5 - 0.5 * math.sqrt(3.0 / 7.0), 0.5, 0.5 + 0.5 * math.sqrt(3.0 / 7.0), 1.0]


def _unit(v):
    v = np.asarray(v, float)
    return v / math.sqrt(float(v @ v))


def coarse_cells():
    """list of (corner directions [v00, v10, v01, v11], r_a index, r_b index) with radii indices into (R0, Rm, R1)."""
    out = []
    for k in range(3):
        i, j = (k + 1) % 3, (k + 2) % 3
        for sj in (-1, 1):
            for si in (-1, 1):
                lo = [0.0, 0.0, 0.0]; lo[i] = si; lo[j] = sj; lo[k] = -1
                hi = list(lo); hi[k] = 1
                ai = [0.0, 0.0, 0.0]; ai[i] = si
                aj = [0.0, 0.0, 0.0]; aj[j] = sj
                v00, v10, v01, v11 = _unit(ai), _unit(lo), _unit(hi), _unit(aj)
                if float(np.cross(v10 - v00, v01 - v00) @ (v00 + v11)) < 0.0:
                    v10, v01 = v01, v10
                blend = lambda x, y: _unit((1 - x) * (1 - y) * v00 + x * (1 - y) * v10 + (1 - x) * y * v01 + x * y * v11)
                for layer in range(2):
                    for sy in range(2):
                        for sx in range(2):
                            corners = [blend(0.5 * (sx + a), 0.5 * (sy + b)) for b in range(2) for a in range(2)]
                            out.append((corners, layer, layer + 1))
    return out


class ShellLevel:
    dim = 3

    def __init__(self, level, inner_radius=0.54, outer_radius=1.0):
        self.level, self.n = level, 1 << level
        self.radii = (inner_radius, 0.5 * (inner_radius + outer_radius), outer_radius)
        self.trees = coarse_cells()
        n, cells = self.n, []
        for t in range(len(self.trees)):
            for m in range(8 ** level):
                ijk = [0, 0, 0]
                for b in range(level):
                    for d in range(3):
                        ijk[d] |= ((m >> (3 * b + d)) & 1) << b
                cells.append((t, tuple(ijk)))
        self.cells = cells
        self.n_cells = len(cells)
        self._number()
        self._geometry()

    def position(self, t, xi):
        c, ia, ib = self.trees[t]
        x, y, z = xi
        p = _unit((1 - x) * (1 - y) * c[0] + x * (1 - y) * c[1] + (1 - x) * y * c[2] + x * y * c[3])
        return (self.radii[ia] + z * (self.radii[ib] - self.radii[ia])) * p

    def _number(self):
        n = self.n
        ord2, ord1 = fe.fe_q_local_order(2, 3), fe.fe_q_local_order(1, 3)
        lex2, lex1 = fe.lex_nodes(3, 3), fe.lex_nodes(2, 3)
        node_u, node_p, xu, xp = {}, {}, [], []

        def key(t, ijk, loc, per):
            x = self.position(t, [(ijk[d] + loc[d] / per) / n for d in range(3)])
            return tuple(int(round((float(v) + 1.0) * 2 ** 19)) for v in x), x

        cu = np.zeros((self.n_cells, 81), dtype=np.int64)
        cp = np.zeros((self.n_cells, 8), dtype=np.int64)
        for m, (t, ijk) in enumerate(self.cells):
            for loc in ord2:
                k, x = key(t, ijk, loc, 2.0)
                if k not in node_u:
                    node_u[k] = len(node_u); xu.append(x)
            for a, loc in enumerate(lex2):
                k, _ = key(t, ijk, loc, 2.0)
                for c in range(3):
                    cu[m, c * 27 + a] = 3 * node_u[k] + c
            for loc in ord1:
                k, x = key(t, ijk, loc, 1.0)
                if k not in node_p:
                    node_p[k] = len(node_p); xp.append(x)
            for a, loc in enumerate(lex1):
                cp[m, a] = node_p[key(t, ijk, loc, 1.0)[0]]
        self.cell_dofs_u, self.cell_dofs_p = cu, cp
        self.n_u, self.n_p = 3 * len(node_u), len(node_p)
        self.node_x_u, self.node_x_p = np.array(xu), np.array(xp)

    def _geometry(self):
        V = np.zeros((3, 5)); D = np.zeros((3, 5))          # 1D Lagrange basis on the Gauss-Lobatto points at the Gauss points
        for q, x in enumerate(fe.GAUSS_X):
            for a in range(5):
                V[q, a] = np.prod([(x - LOBATTO5[b]) / (LOBATTO5[a] - LOBATTO5[b]) for b in range(5) if b != a])
                D[q, a] = sum(np.prod([(x - LOBATTO5[c]) / (LOBATTO5[a] - LOBATTO5[c]) for c in range(5) if c not in (a, b)])
                              / (LOBATTO5[a] - LOBATTO5[b]) for b in range(5) if b != a)
        qs = fe.lex_nodes(3, 3)
        ks = fe.lex_nodes(5, 3)
        N = np.array([[V[q[0], k[0]] * V[q[1], k[1]] * V[q[2], k[2]] for k in ks] for q in qs])
        dN = np.array([[[np.prod([(D if d == r else V)[q[d], k[d]] for d in range(3)]) for r in range(3)] for k in ks] for q in qs])
        w = np.array([fe.GAUSS_W[q[0]] * fe.GAUSS_W[q[1]] * fe.GAUSS_W[q[2]] for q in qs])
        n = self.n
        self.xq = np.zeros((self.n_cells, 27, 3)); self.jinv = np.zeros((self.n_cells, 27, 3, 3)); self.jxw = np.zeros((self.n_cells, 27))
        for m, (t, ijk) in enumerate(self.cells):
            X = np.array([self.position(t, [(ijk[d] + LOBATTO5[k[d]]) / n for d in range(3)]) for k in ks])    # (125, 3)
            self.xq[m] = N @ X
            for q in range(27):
                J = X.T @ dN[q]                                       # J[d][r] = d x_d / d xihat_r
                self.jinv[m, q] = np.linalg.inv(J)
                self.jxw[m, q] = np.linalg.det(J) * w[q]

    def quadrature_points(self):
        return self.xq

    def midpoint_geometry(self):
        """(position, det J) at the cell midpoints in the degree-4 mapping (QMidpoint, helper_functions.cc:286)."""
        ks = fe.lex_nodes(5, 3)
        L = lambda a, x: np.prod([(x - LOBATTO5[b]) / (LOBATTO5[a] - LOBATTO5[b]) for b in range(5) if b != a])
        dL = lambda a, x: sum(np.prod([(x - LOBATTO5[c]) / (LOBATTO5[a] - LOBATTO5[c]) for c in range(5) if c not in (a, b)])
                              / (LOBATTO5[a] - LOBATTO5[b]) for b in range(5) if b != a)
        v = [L(a, 0.5) for a in range(5)]; d = [dL(a, 0.5) for a in range(5)]
        N = np.array([v[k[0]] * v[k[1]] * v[k[2]] for k in ks])
        dN = np.array([[np.prod([(d if e == r else v)[k[e]] for e in range(3)]) for r in range(3)] for k in ks])
        xm = np.zeros((self.n_cells, 3)); det = np.zeros(self.n_cells)
        for m, (t, ijk) in enumerate(self.cells):
            X = np.array([self.position(t, [(ijk[e] + LOBATTO5[k[e]]) / self.n for e in range(3)]) for k in ks])
            xm[m] = N @ X
            det[m] = np.linalg.det(X.T @ dN)
        return xm, det

    def constrained_u(self, bc):
        if bc == "none":
            return np.zeros(0, dtype=np.int64)
        assert bc == "zero"
        out = []
        for k, x in enumerate(self.node_x_u):
            r = math.sqrt(float(x @ x))
            if abs(r - self.radii[0]) < 1e-9 or abs(r - self.radii[2]) < 1e-9:
                out += [3 * k, 3 * k + 1, 3 * k + 2]
        return np.array(out, dtype=np.int64)


class ShellHierarchy:
    def __init__(self, refinement, inner_radius=0.54, outer_radius=1.0):
        self.dim, self.R = 3, refinement
        self.levels = [ShellLevel(l, inner_radius, outer_radius) for l in range(refinement + 1)]
        self.active = self.levels[-1]

    def children(self, level):
        npar = self.levels[level - 1].n_cells
        return (np.arange(npar)[:, None] * 8 + np.arange(8)[None, :]).astype(np.int64)
