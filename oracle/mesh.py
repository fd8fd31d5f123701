"""Box mesh hierarchy, DoF numbering and constraint sets (oracle; test infrastructure).

Literal, loop-based restatement of the rules in SURVEY.md App. A.4 for the meshes
ASPECT's box geometry produces (/root/reference/source/geometry_model/box.cc:66-74,
one coarse cell, `Initial global refinement = R`):

* level-l cells are the 2^(dim*l) children in Morton (z-) order, children ordered
  x fastest (deal.II child numbering);
* DoFs are numbered in first-touch order while visiting the cells in that order and,
  inside a cell, the FE-local order vertices -> lines -> quads -> hex interior with the
  vector components innermost (FESystem(FE_Q(2),dim)); this is what
  `DoFRenumbering::hierarchical` produces on the active mesh
  (stokes_matrix_free_local_smoothing.cc:1488-1491, :1559-1561) and what
  `distribute_mg_dofs` produces per level (:1594-1596);
* MatrixFree-style cell->DoF maps are lexicographic in the node and component-major
  ([c0: 3^dim][c1: 3^dim]...).
Constraint sets follow :1501-1553 / :1598-1633 for the two boundary conditions the
benchmark configs use on a box: "zero velocity" (all components on every face) and
"tangential velocity" (= no normal flux: component d on the faces normal to d).
"""
import numpy as np
from . import fe


def morton_cells(level, dim):
    """Integer cell coordinates (n_cells, dim) of level `level` in Morton order."""
    n_cells = 1 << (dim * level)
    out = np.zeros((n_cells, dim), dtype=np.int64)
    for m in range(n_cells):
        for b in range(level):
            for d in range(dim):
                out[m, d] |= ((m >> (dim * b + d)) & 1) << b
    return out


class Level:
    """One level of the hierarchy (also used for the active mesh)."""

    def __init__(self, dim, level, extents=None):
        self.dim = dim
        self.level = level
        self.n = n = 1 << level
        self.extents = np.ones(dim) if extents is None else np.asarray(extents, float)
        self.h = self.extents / n
        self.cells = morton_cells(level, dim)
        self.n_cells = len(self.cells)
        self._number()

    def _number(self):
        dim, n = self.dim, self.n
        g2 = 2 * n + 1          # Q2 node grid
        g1 = n + 1              # Q1 node grid
        ord2 = fe.fe_q_local_order(2, dim)
        ord1 = fe.fe_q_local_order(1, dim)
        lex2 = fe.lex_nodes(3, dim)
        lex1 = fe.lex_nodes(2, dim)
        node_u, node_p = {}, {}
        cu = np.zeros((self.n_cells, dim * 3 ** dim), dtype=np.int64)
        cp = np.zeros((self.n_cells, 2 ** dim), dtype=np.int64)
        for m, ijk in enumerate(self.cells):
            # velocity: first touch in FE-local order; components of a node are consecutive
            for loc in ord2:
                key = tuple(2 * int(ijk[d]) + loc[d] for d in range(dim))
                if key not in node_u:
                    node_u[key] = len(node_u)
            for a, loc in enumerate(lex2):
                k = node_u[tuple(2 * int(ijk[d]) + loc[d] for d in range(dim))]
                for c in range(dim):
                    cu[m, c * 3 ** dim + a] = dim * k + c
            for loc in ord1:
                key = tuple(int(ijk[d]) + loc[d] for d in range(dim))
                if key not in node_p:
                    node_p[key] = len(node_p)
            for a, loc in enumerate(lex1):
                cp[m, a] = node_p[tuple(int(ijk[d]) + loc[d] for d in range(dim))]
        assert len(node_u) == g2 ** dim and len(node_p) == g1 ** dim
        self.cell_dofs_u = cu
        self.cell_dofs_p = cp
        self.n_u = dim * len(node_u)
        self.n_p = len(node_p)
        # node number -> integer grid coordinate
        self.node_grid_u = np.zeros((len(node_u), dim), dtype=np.int64)
        for key, k in node_u.items():
            self.node_grid_u[k] = key
        self.node_grid_p = np.zeros((len(node_p), dim), dtype=np.int64)
        for key, k in node_p.items():
            self.node_grid_p[k] = key

    # -- geometry helpers -------------------------------------------------
    def dof_coords_u(self):
        """(n_u, dim) coordinates of the support point of every velocity DoF."""
        x = self.node_grid_u * (self.h / 2.0)
        return np.repeat(x, self.dim, axis=0)

    def dof_coords_p(self):
        return self.node_grid_p * self.h

    def quadrature_points(self):
        """(n_cells, 3^dim, dim) physical Gauss points, q lexicographic."""
        t = fe.tensor_tables(self.dim)
        return (self.cells[:, None, :] + t["xq"][None, :, :]) * self.h

    def cell_volume(self):
        return float(np.prod(self.h))

    # -- constraints --------------------------------------------------------
    def constrained_u(self, bc):
        """Sorted constrained velocity DoFs for bc in {'zero', 'tangential', 'none'}."""
        if bc == "none":
            return np.zeros(0, dtype=np.int64)
        g = self.node_grid_u
        hi = 2 * self.n
        out = []
        for k in range(len(g)):
            for c in range(self.dim):
                if bc == "zero":
                    on = any(g[k, d] == 0 or g[k, d] == hi for d in range(self.dim))
                elif bc == "tangential":
                    on = g[k, c] == 0 or g[k, c] == hi
                else:
                    raise ValueError(bc)
                if on:
                    out.append(self.dim * k + c)
        return np.array(sorted(out), dtype=np.int64)


class Hierarchy:
    """Levels 0..R of a uniformly refined box; the active mesh is level R
    (its hierarchical numbering coincides with the level numbering, see module doc)."""

    def __init__(self, dim, refinement, extents=None):
        self.dim = dim
        self.R = refinement
        self.levels = [Level(dim, l, extents) for l in range(refinement + 1)]
        self.active = self.levels[-1]

    def children(self, level):
        """(n_cells(level-1), 2^dim) indices of the children (level `level`) of every
        level-1 cell; Morton order makes them 2^dim*parent + child."""
        npar = self.levels[level - 1].n_cells
        return (np.arange(npar)[:, None] * (1 << self.dim) + np.arange(1 << self.dim)[None, :]).astype(np.int64)
