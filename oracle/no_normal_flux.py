"""TEST INFRASTRUCTURE (oracle): no-normal-flux (`Tangential velocity boundary indicators`) on CURVED boundaries, CPU only.

The product rejects this boundary condition on curved cells; this module restates what the reference's
LOCAL-SMOOTHING handler does with it, on assembled matrices, so that the one local-smoothing golden the reference holds for the
96-cell shell mesh -- tests/shell_3d_gmg/screen-output, first solve: 24+0 iterations -- can be checked against the same
restatement of smoothers, transfers and Krylov solver that every other test uses (oracle/solver.py).

Reference:
* active level: `VectorTools::compute_no_normal_flux_constraints(..., use_manifold_for_normal = true)`
  (source/simulator/core.cc, compute_initial_velocity_boundary_constraints): at every boundary node the component with the largest
  |n_k| (ties: the first) is constrained, u_k = - sum_{j != k} (n_j / n_k) u_j, n = the manifold's normal = the radial direction;
* levels: `compute_no_normal_flux_constraints_on_level` into user constraints of `MGConstrainedDoFs`, merged into the level
  constraints of the level MatrixFree objects, Dirichlet values win (stokes_matrix_free_local_smoothing.cc:1719-1789);
* a MatrixFree operator with constraints C is C^T A C on the unconstrained DoFs and the identity on the constrained ones
  (read_dof_values resolves, distribute_local_to_global condenses, MatrixFreeOperators::Base adds the unit diagonal);
  `MatrixFreeTools::compute_diagonal` returns that operator's diagonal (matrix_free_operators.cc:1057-1093, :879);
* `MGTransferMatrixFree::prolongate` distributes the user constraints of the coarse level on its source vector before it
  interpolates; `restrict_and_add` is the plain transpose; Dirichlet level DoFs are zeroed on the coarse side in both
  (deal.II 9.6 mg_transfer_matrix_free.cc, recalled).
* right-hand side: `distribute_local_to_global` = C^T f; solution: `constraints.distribute` = C u.

Gravity of that test is `ascii data` with the default profile data/gravity-model/prem.txt (44 depth / value pairs, a constant
table reproduced below), linear in depth (source/gravity_model/ascii_data.cc); material `simple` (source/material_model/simple.cc):
constant viscosity, rho = rho_0 (1 - alpha (T - T_0)).
"""
import numpy as np
import scipy.sparse as sp

from . import assembled, fe, solver

# data/gravity-model/prem.txt: depth [m], gravity [m/s^2]
PREM = np.array([
    [0, 9.816], [3000, 9.822], [15000, 9.833], [24000, 9.839], [80000, 9.855], [150000, 9.878], [220000, 9.905], [310000, 9.936],
    [400000, 9.969], [500000, 9.988], [600000, 10.004], [670000, 10.014], [771000, 9.999], [971000, 9.97], [1171000, 9.947],
    [1371000, 9.933], [1571000, 9.931], [1771000, 9.947], [1971000, 9.986], [2171000, 10.054], [2371000, 10.158], [2571000, 10.31],
    [2741000, 10.484], [2771000, 10.52], [2891000, 10.682], [2971000, 10.507], [3171000, 10.046], [3371000, 9.557], [3571000, 9.041],
    [3771000, 8.502], [3971000, 7.943], [4171000, 7.365], [4371000, 6.772], [4571000, 6.167], [4771000, 5.555], [4971000, 4.941],
    [5149000, 4.4], [5171000, 4.325], [5371000, 3.62], [5571000, 2.907], [5771000, 2.186], [5971000, 1.46], [6171000, 0.731],
    [6371000, 0.0]])


def prem_gravity(depth):
    return np.interp(depth, PREM[:, 0], PREM[:, 1])


def boundary_nodes(level, radii, rtol=1e-9):
    """Nodes of the Q2 lattice on the spheres of the given radius / radii (None or (): no nodes)."""
    if radii is None:
        return np.zeros(0, dtype=np.int64)
    r = np.sqrt((level.node_x_u ** 2).sum(axis=1))
    hit = np.zeros(len(r), bool)
    for radius in np.atleast_1d(radii):
        hit |= np.abs(r - radius) < rtol * radius
    return np.nonzero(hit)[0]


def constraint_matrix(level, dirichlet_nodes, flux_nodes):
    """C (n_u x n_u): identity rows for free DoFs, zero rows for Dirichlet DoFs, u_k = -sum (n_j / n_k) u_j for the constrained
    component k of a no-normal-flux node.  Returns (C, constrained DoF indices)."""
    n = level.n_u
    rows, cols, vals = [], [], []
    cons = np.zeros(n, bool)
    for k in dirichlet_nodes:
        cons[3 * k:3 * k + 3] = True
    for k in flux_nodes:
        x = level.node_x_u[k]
        nrm = x / np.sqrt(x @ x)
        a = np.abs(nrm)
        c = int(np.argmax(a >= a.max() - 1e-9))                       # the first component within 1e-9 of the largest
        cons[3 * k + c] = True
        for j in range(3):
            if j != c:
                rows.append(3 * k + c); cols.append(3 * k + j); vals.append(-nrm[j] / nrm[c])
    free = np.nonzero(~cons)[0]
    rows = np.concatenate([free, np.array(rows, dtype=np.int64)])
    cols = np.concatenate([free, np.array(cols, dtype=np.int64)])
    vals = np.concatenate([np.ones(len(free)), np.array(vals, dtype=float)])
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr(), np.nonzero(cons)[0]


def _constrain(A, C, cons):
    Ac = (C.T @ A @ C).tolil()
    for i in cons:
        Ac[i, i] = 1.0
    return Ac.tocsr()


class NoNormalFluxBackend(solver.AssembledBackend):
    """AssembledBackend with zero velocity on the nodes of radius `r_zero` and no normal flux on those of radius `r_flux`
    (each a radius, a tuple of radii or None)."""

    def __init__(self, hier, active_eta, level_eta, pressure_scaling, r_zero, r_flux):
        self.hier = hier
        self.n_levels = hier.R + 1
        act = hier.active
        CC = [constraint_matrix(lev, boundary_nodes(lev, r_zero), boundary_nodes(lev, r_flux)) for lev in hier.levels]
        self.C = [c for c, _ in CC]
        self.cu = [k for _, k in CC]
        Ca, ka = CC[-1]
        K, A, BT = assembled.stokes_matrix(act, active_eta, pressure_scaling, None)
        self.A_act = _constrain(A, Ca, ka)
        self.BT_act = (Ca.T @ BT).tocsr()
        self.K = sp.bmat([[self.A_act, self.BT_act], [self.BT_act.T, None]], format="csr")
        self.M_act = assembled.assemble_mass(act, active_eta, pressure_scaling)
        self.A = [_constrain(assembled.assemble_A(lev, level_eta[l], None), self.C[l], self.cu[l]) for l, lev in enumerate(hier.levels)]
        self.M = [assembled.assemble_mass(lev, level_eta[l], pressure_scaling) for l, lev in enumerate(hier.levels)]
        plain = [None] + [solver.prolongation_matrix(hier, l, "u", None) for l in range(1, self.n_levels)]
        # prolongation: distribute (and zero the Dirichlet DoFs) on the coarse side, then interpolate
        self.Pu = [None] + [(plain[l] @ self.C[l - 1]).tocsr() for l in range(1, self.n_levels)]
        # restriction: plain transpose, Dirichlet DoFs of the coarse level zeroed
        self.Ru = [None]
        for l in range(1, self.n_levels):
            lev = hier.levels[l - 1]
            keep = np.ones(lev.n_u)
            for k in boundary_nodes(lev, r_zero):
                keep[3 * k:3 * k + 3] = 0.0
            self.Ru.append((sp.diags(keep) @ plain[l].T).tocsr())
        self.Pp = [None] + [solver.prolongation_matrix(hier, l, "p") for l in range(1, self.n_levels)]
        self.n_u, self.n_p = act.n_u, act.n_p

    def restrict_u(self, l, x): return self.Ru[l] @ x

    def condense_rhs(self, f):
        """distribute_local_to_global of the cell right-hand sides: C^T f, constrained rows zero."""
        g = self.C[-1].T @ f
        g[self.cu[-1]] = 0.0
        return g

    def distribute(self, u):
        return self.C[-1] @ u


def shell_3d_gmg_problem(hier, r_inner=3481000.0, r_outer=6336000.0):
    """Coefficients of tests/shell_3d_gmg.prm on the given shell hierarchy: viscosity table (constant 1e22), force density
    rho g at the quadrature points, pressure scaling.  Temperature: 1473 K with the boundary values (973 K outer, 1973 K inner)
    in the boundary nodes of the Q2 field (initial_conditions.cc:232-233), interpolated to the quadrature points."""
    L = hier.active
    t = fe.tensor_tables(3)
    T_node = np.full(len(L.node_x_u), 1473.0)
    T_node[boundary_nodes(L, r_outer)] = 973.0
    T_node[boundary_nodes(L, r_inner)] = 1973.0
    cell_nodes = L.cell_dofs_u[:, :27] // 3
    T_q = T_node[cell_nodes] @ t["phi2"].T                            # (cells, nq)
    rho_q = 3300.0 * (1.0 - 4e-5 * (T_q - 293.0))
    r_q = np.sqrt((L.xq ** 2).sum(axis=-1))
    g_q = -prem_gravity(r_outer - r_q)[..., None] * L.xq / r_q[..., None]
    eta_q = np.full(L.xq.shape[:2], 1e22)
    ps = 1e22 / (1e4 / (6336000.0 - 3481000.0) * (r_outer - r_inner))
    return eta_q, rho_q[..., None] * g_q, ps
