"""Restatement of the block-GMG Stokes solve (oracle; test infrastructure).

Follows /root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:758-1313
(solve(): smoother set-up :774-825, eigenvalue estimation :834-849, coarse solver :855-860,
tolerance :977-1041, cheap GMRES :1277-1301),
/root/reference/include/aspect/simulator/solver/block_stokes_preconditioner.h:220-255
and the deal.II 9.6 semantics recalled in SURVEY.md App. A (PreconditionChebyshev A.2,
Multigrid/PreconditionMG A.3, MGTransferMatrixFree A.3, SolverGMRES A.5).

The solver is written against a small "backend" interface so the same driver runs over
assembled sparse matrices (AssembledBackend, below) or over the C matrix-free kernels
(oracle/cpu.py):
    n_levels, apply_A(l, x), apply_M(l, x), dinv_A(l), dinv_M(l),
    prolongate_u/p(l, xc) -> xf (level l-1 -> l), restrict_u/p(l, xf) -> xc,
    stokes(xu, xp) -> (yu, yp), BT(xp) -> yu        (active level)
"""
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
from . import fe, assembled


# ---------------------------------------------------------------------------
# transfer matrices (MGTransferMatrixFree semantics, uniform refinement)
# ---------------------------------------------------------------------------
def prolongation_matrix(hier, l, space, constrained_coarse=None):
    """Sparse P (n_fine x n_coarse) from level l-1 to level l for space 'u' or 'p'.
    Per parent cell: tensor-product 1D prolongation onto the child patch, weighted by
    1/valence and added (App. A.3); coarse Dirichlet DoFs are zeroed before prolongation."""
    dim = hier.dim
    fine, coarse = hier.levels[l], hier.levels[l - 1]
    if space == "u":
        P1, npd, ncomp = fe.P2_1D, 3, dim
        cf, cc, nf, nc = fine.cell_dofs_u, coarse.cell_dofs_u, fine.n_u, coarse.n_u
    else:
        P1, npd, ncomp = fe.P1_1D, 2, 1
        cf, cc, nf, nc = fine.cell_dofs_p, coarse.cell_dofs_p, fine.n_p, coarse.n_p
    npatch = 2 * npd - 1
    Ploc = P1
    for _ in range(dim - 1):
        Ploc = np.kron(P1, Ploc)                      # lexicographic, x fastest
    nn_c, nn_f = npd ** dim, npd ** dim
    patch_nodes = fe.lex_nodes(npatch, dim)
    child_nodes = fe.lex_nodes(npd, dim)
    ch = hier.children(l)
    rows, cols, vals = [], [], []
    for c in range(ncomp):
        # patch fine dof indices per parent: fill from the children
        patch = np.zeros((coarse.n_cells, npatch ** dim), dtype=np.int64)
        for k in range(1 << dim):
            off = [(k >> d) & 1 for d in range(dim)]
            for a, loc in enumerate(child_nodes):
                pn = tuple(loc[d] + (npd - 1) * off[d] for d in range(dim))
                pi = sum(pn[d] * npatch ** d for d in range(dim))
                patch[:, pi] = cf[ch[:, k], c * nn_f + a]
        cdofs = cc[:, c * nn_c:(c + 1) * nn_c]
        R = np.repeat(patch[:, :, None], nn_c, axis=2)
        C = np.repeat(cdofs[:, None, :], npatch ** dim, axis=1)
        V = np.broadcast_to(Ploc[None], R.shape)
        rows.append(R.ravel()); cols.append(C.ravel()); vals.append(V.ravel())
    rows, cols, vals = map(np.concatenate, (rows, cols, vals))
    # 1/valence weights on the fine side
    valence = np.bincount(rows, minlength=nf).astype(float) / nn_c
    vals = vals / valence[rows]
    P = sp.coo_matrix((vals, (rows, cols)), shape=(nf, nc)).tocsr()
    if constrained_coarse is not None and len(constrained_coarse):
        keep = np.ones(nc)
        keep[constrained_coarse] = 0.0
        P = P @ sp.diags(keep)
    P.eliminate_zeros()
    return P


# ---------------------------------------------------------------------------
class AssembledBackend:
    """Backend over assembled sparse matrices."""

    def __init__(self, hier, active_eta, level_eta, pressure_scaling, bc, compressible=False):
        self.hier = hier
        self.n_levels = hier.R + 1
        act = hier.active
        self.cu = [lev.constrained_u(bc) for lev in hier.levels]
        self.K, self.A_act, self.BT_act = assembled.stokes_matrix(act, active_eta, pressure_scaling, self.cu[-1], compressible)
        self.M_act = assembled.assemble_mass(act, active_eta, pressure_scaling)
        self.A = [assembled.assemble_A(lev, level_eta[l], self.cu[l], compressible) for l, lev in enumerate(hier.levels)]
        self.M = [assembled.assemble_mass(lev, level_eta[l], pressure_scaling) for l, lev in enumerate(hier.levels)]
        self.Pu = [None] + [prolongation_matrix(hier, l, "u", self.cu[l - 1]) for l in range(1, self.n_levels)]
        self.Pp = [None] + [prolongation_matrix(hier, l, "p") for l in range(1, self.n_levels)]
        self.n_u, self.n_p = act.n_u, act.n_p

    def apply_A(self, l, x): return self.A[l] @ x
    def apply_M(self, l, x): return self.M[l] @ x
    def dinv_A(self, l): return 1.0 / self.A[l].diagonal()
    def dinv_M(self, l): return 1.0 / self.M[l].diagonal()
    def prolongate_u(self, l, x): return self.Pu[l] @ x
    def prolongate_p(self, l, x): return self.Pp[l] @ x
    def restrict_u(self, l, x): return self.Pu[l].T @ x
    def restrict_p(self, l, x): return self.Pp[l].T @ x

    def stokes(self, xu, xp):
        y = self.K @ np.concatenate([xu, xp])
        return y[:self.n_u], y[self.n_u:]

    def BT(self, xp): return self.BT_act @ xp
    def A_active(self, xu): return self.A_act @ xu
    def M_active(self, xp): return self.M_act @ xp


# ---------------------------------------------------------------------------
# PreconditionChebyshev (deal.II 9.6, first kind), App. A.2
# ---------------------------------------------------------------------------
class Chebyshev:
    def __init__(self, apply, dinv, degree, smoothing_range, eig_cg_n_iterations, eig_cg_residual=1e-7):
        self.apply, self.dinv, self.degree = apply, dinv, degree
        self.smoothing_range = smoothing_range
        self.eig_its, self.eig_res = eig_cg_n_iterations, eig_cg_residual
        self.n_matvec = 0

    def estimate_eigenvalues(self):
        n = len(self.dinv)
        b = (np.arange(n) % 11).astype(float)          # set_initial_guess: global index mod 11
        b -= b.mean()
        lam_min, lam_max, its = _cg_lanczos(self.apply, self.dinv, b, self.eig_its, self.eig_res)
        self.n_matvec += its
        self.cg_iterations = its
        self.min_eig_est, self.max_eig_est = lam_min, lam_max
        lam_max = 1.2 * lam_max
        if self.smoothing_range > 1.0:
            alpha = lam_max / self.smoothing_range
        else:
            alpha = min(0.9 * lam_max, lam_min)
        self.delta = 0.5 * (lam_max - alpha)
        self.theta = 0.5 * (lam_max + alpha)

    def _mv(self, x):
        self.n_matvec += 1
        return self.apply(x)

    def vmult(self, b):
        """x = p(A) b from a zero initial guess: degree-1 matvecs."""
        x = self.dinv * b / self.theta
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return x
        x_old = np.zeros_like(x)
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for _ in range(self.degree - 1):
            t = self._mv(x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            x, x_old = (1.0 + f1) * x - f1 * x_old + f2 * self.dinv * (b - t), x
        return x

    def step(self, x, b):
        """one sweep from a non-zero guess: degree matvecs."""
        t = self._mv(x)
        x, x_old = x + self.dinv * (b - t) / self.theta, x
        if self.degree < 2 or abs(self.delta) < 1e-40:
            return x
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        for _ in range(self.degree - 1):
            t = self._mv(x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2 = rhokp * rhok, 2.0 * rhokp / self.delta
            rhok = rhokp
            x, x_old = (1.0 + f1) * x - f1 * x_old + f2 * self.dinv * (b - t), x
        return x


def _cg_lanczos(apply, dinv, b, max_its, tol):
    """Jacobi-preconditioned CG from x=0 (deal.II SolverCG with IterationNumberControl);
    returns (lambda_min, lambda_max, iterations) of the Lanczos tridiagonal."""
    x = np.zeros_like(b)
    r = b.copy()
    res = np.linalg.norm(r)
    diag, off = [], []
    it = 0
    if res <= tol or max_its == 0:
        return 1.0, 1.0, 0
    p = None
    rz_old = alpha_old = beta = None
    while True:
        it += 1
        z = dinv * r
        rz = r @ z
        if it > 1:
            beta = rz / rz_old
            p = z + beta * p
        else:
            p = z.copy()
        v = apply(p)
        alpha = rz / (p @ v)
        x += alpha * p
        r -= alpha * v
        res = np.linalg.norm(r)
        if it > 1:
            diag.append(1.0 / alpha + beta / alpha_old)
            off.append(np.sqrt(beta) / alpha_old)
        else:
            diag.append(1.0 / alpha)
        rz_old, alpha_old = rz, alpha
        if it >= max_its or res <= tol:
            break
    ev = sla.eigvalsh_tridiagonal(np.array(diag), np.array(off)) if len(diag) > 1 else np.array(diag)
    return float(ev.min()), float(ev.max()), it


# ---------------------------------------------------------------------------
# Multigrid V-cycle with SmootherRelaxation semantics, App. A.3
# ---------------------------------------------------------------------------
class VCycle:
    def __init__(self, n_levels, apply, dinv, prolongate, restrict, smoother_steps):
        self.n_levels, self.apply, self.prolongate, self.restrict = n_levels, apply, prolongate, restrict
        self.steps = smoother_steps
        self.smoothers = []
        for l in range(n_levels):
            if l > 0:
                s = Chebyshev(lambda x, l=l: apply(l, x), dinv(l), 4, 15.0, 10)
            else:
                s = Chebyshev(lambda x: apply(0, x), dinv(0), 8, 1e-3, 100)
            s.estimate_eigenvalues()
            self.smoothers.append(s)

    def _apply_smoother(self, l, b):          # SmootherRelaxation::apply
        x = self.smoothers[l].vmult(b)
        for _ in range(1, self.steps):
            x = self.smoothers[l].step(x, b)
        return x

    def _smooth(self, l, x, b):               # SmootherRelaxation::smooth
        for _ in range(self.steps):
            x = self.smoothers[l].step(x, b)
        return x

    def _level(self, l, defect):
        if l == 0:
            return self._apply_smoother(0, defect)          # MGCoarseGridApplySmoother
        x = self._apply_smoother(l, defect)
        t = defect - self.apply(l, x)
        xc = self._level(l - 1, self.restrict(l, t))
        x = x + self.prolongate(l, xc)
        return self._smooth(l, x, defect)

    def vmult(self, src):
        return self._level(self.n_levels - 1, src)


# ---------------------------------------------------------------------------
# SolverGMRES (right preconditioning, delayed classical Gram-Schmidt), App. A.5
# ---------------------------------------------------------------------------
# deal.II 9.6's default OrthogonalizationStrategy is delayed_classical_gram_schmidt = classical Gram-Schmidt
# with TWO orthogonalisation passes whose second pass is delayed by one iteration to share a reduction
# (DCGS2, Bielich et al. 2022).  Numerically it behaves like CGS2 (orthogonality at machine precision), which
# is what is restated here: classical Gram-Schmidt, always re-orthogonalised.  With single-pass CGS the
# sol_cx goldens (pressure scaling 1e5) are not reproducible: the Arnoldi basis loses orthogonality and
# the residual stagnates just above the tolerance (11 -> 13..21 iterations depending on round-off).
REORTHOGONALIZE_ALWAYS = True


def gmres_right(op, prec, b, x0, tol, restart, max_its, history=None):
    """Returns (x, iterations, last_residual, converged).  `restart` = basis size."""
    x = x0.copy()
    its = 0
    eps = np.finfo(float).eps
    re_orth = REORTHOGONALIZE_ALWAYS
    state = "iterate"
    rho = np.nan
    while state == "iterate":
        p = b - op(x)
        rho = np.linalg.norm(p)
        if history is not None and its == 0:
            history.append(rho)
        if rho <= tol:
            state = "success"; break
        if its >= max_its:
            state = "failure"; break
        V = [p / rho]
        gamma = np.zeros(restart + 1); gamma[0] = rho
        H = np.zeros((restart + 1, restart))
        ci, si = np.zeros(restart), np.zeros(restart)
        dim = 0
        for j in range(restart):
            if state != "iterate":
                break
            its += 1
            vv = op(prec(V[j]))
            dim = j + 1
            h = np.zeros(dim + 1)
            consider = (not re_orth) and (its % 5 == 0)
            if consider:
                norm_start = np.linalg.norm(vv)
            Vm = np.stack(V[:dim], axis=1)
            hh = Vm.T @ vv
            vv = vv - Vm @ hh
            h[:dim] = hh
            nrm = np.linalg.norm(vv)
            if consider and not (nrm > 10.0 * norm_start * np.sqrt(eps)):
                re_orth = True
            if re_orth:
                hh = Vm.T @ vv
                vv = vv - Vm @ hh
                h[:dim] += hh
                nrm = np.linalg.norm(vv)
            h[dim] = nrm
            if nrm != 0:
                vv = vv / nrm
            V.append(vv)
            # Givens
            for i in range(j):
                t = h[i]
                h[i] = ci[i] * t + si[i] * h[i + 1]
                h[i + 1] = -si[i] * t + ci[i] * h[i + 1]
            s = np.hypot(h[j], h[j + 1])
            ci[j], si[j] = h[j] / s, h[j + 1] / s
            h[j] = s
            gamma[j + 1] = -si[j] * gamma[j]
            gamma[j] *= ci[j]
            H[:dim, j] = h[:dim]
            rho = abs(gamma[dim])
            if history is not None:
                history.append(rho)
            if rho <= tol:
                state = "success"
            elif its >= max_its:
                state = "failure"
        y = sla.solve_triangular(H[:dim, :dim], gamma[:dim])
        pp = np.zeros_like(b)
        for i in range(dim):
            pp += y[i] * V[i]
        x = x + prec(pp)
    return x, its, rho, state == "success"


def cg(apply, prec, b, tol, max_its):
    """deal.II SolverCG from a zero start vector (InverseVelocityBlock / SchurApproximation with do_solve = true,
    block_stokes_preconditioner.h:108-135, :312-345).  Returns (x, iterations = SolverControl::last_step, converged)."""
    x = np.zeros_like(b)
    g = -b                                  # g = A x - b
    res = np.linalg.norm(g)
    if res <= tol:
        return x, 0, True
    d = None
    gh = 0.0
    it = 0
    while True:
        it += 1
        h = prec(g)
        if it > 1:
            beta = gh
            gh = g @ h
            beta = gh / beta
            d = -h + beta * d
        else:
            d = -h
            gh = g @ h
        Ad = apply(d)
        alpha = gh / (d @ Ad)
        x = x + alpha * d
        g = g + alpha * Ad
        res = np.linalg.norm(g)
        if res <= tol:
            return x, it, True
        if it >= max_its:
            return x, it, False


def bicgstab(apply, prec, b, tol, max_its, breakdown=1e-10):
    """deal.II 9.6 SolverBicgstab(AdditionalData(exact_residual=False)) from a zero start vector (InverseVelocityBlock::vmult with
    a non-symmetric A block, block_stokes_preconditioner.h:119-132), recalled from solver_bicgstab.h: iterate() with the early
    success check after the first half step (x += alpha y), residual norms from add_and_dot, restart after a breakdown.
    Returns (x, iterations = SolverControl::last_step, converged)."""
    x = np.zeros_like(b)
    step = 0
    while True:
        r = b - apply(x)
        rbar = r.copy()
        res = np.linalg.norm(r)
        if res <= tol:
            return x, step, True
        if step >= max_its:
            return x, step, False
        alpha = rho = omega = 1.0
        first = step + 1
        broke = False
        p = v = None
        while True:
            step += 1
            rhobar = res * res if step == first else r @ rbar
            if abs(rhobar) < breakdown:
                broke = True; break
            beta = rhobar * alpha / (rho * omega)
            rho = rhobar
            p = r.copy() if step == first else beta * p + r - beta * omega * v
            y = prec(p)
            v = apply(y)
            rbar_v = rbar @ v
            if abs(rbar_v) < breakdown:
                broke = True; break
            alpha = rho / rbar_v
            r = r - alpha * v
            res = np.linalg.norm(r)
            if res <= tol:
                return x + alpha * y, step, True
            z = prec(r)
            t = apply(z)
            t_r, t_t = t @ r, t @ t
            if t_t < breakdown:
                broke = True; break
            omega = t_r / t_t
            x = x + alpha * y + omega * z
            r = r - omega * t
            res = np.linalg.norm(r)
            if res <= tol:
                return x, step, True
            if step >= max_its:
                return x, step, False
        if not broke:
            return x, step, True


def fgmres(op, prec, b, x0, tol, restart, max_its, history=None):
    """deal.II SolverFGMRES(AdditionalData(max_basis_size = restart)) (local_smoothing.cc:1333-1350): flexible GMRES,
    the preconditioned vectors z_j are stored and x += Z y.  Returns (x, iterations, last_residual, converged)."""
    x = x0.copy()
    its = 0
    state = "iterate"
    rho = np.nan
    while state == "iterate":
        p = b - op(x)
        rho = np.linalg.norm(p)
        if history is not None and its == 0:
            history.append(rho)
        if rho <= tol:
            state = "success"; break
        if its >= max_its:
            state = "failure"; break
        V, Z = [p / rho], []
        gamma = np.zeros(restart + 1); gamma[0] = rho
        H = np.zeros((restart + 1, restart))
        ci, si = np.zeros(restart), np.zeros(restart)
        dim = 0
        for j in range(restart):
            if state != "iterate":
                break
            its += 1
            Z.append(prec(V[j]))
            vv = op(Z[j])
            dim = j + 1
            Vm = np.stack(V[:dim], axis=1)
            h = np.zeros(dim + 1)
            for _ in range(2):                   # CGS2 (= deal.II's delayed classical Gram-Schmidt numerics)
                hh = Vm.T @ vv
                vv = vv - Vm @ hh
                h[:dim] += hh
            h[dim] = np.linalg.norm(vv)
            V.append(vv / h[dim] if h[dim] != 0 else vv)
            for i in range(j):
                t = h[i]
                h[i] = ci[i] * t + si[i] * h[i + 1]
                h[i + 1] = -si[i] * t + ci[i] * h[i + 1]
            s = np.hypot(h[j], h[j + 1])
            ci[j], si[j] = h[j] / s, h[j + 1] / s
            h[j] = s
            gamma[j + 1] = -si[j] * gamma[j]
            gamma[j] *= ci[j]
            H[:dim, j] = h[:dim]
            rho = abs(gamma[dim])
            if history is not None:
                history.append(rho)
            if rho <= tol:
                state = "success"
            elif its >= max_its:
                state = "failure"
        y = sla.solve_triangular(H[:dim, :dim], gamma[:dim])
        for i in range(dim):
            x = x + y[i] * Z[i]
    return x, its, rho, state == "success"


def libstdcxx_normal_stream(n):
    """The first n values of `std::mt19937 rng; std::normal_distribution<double> nd(0, 1); nd(rng)` as libstdc++
    produces them (default seed 5489; generate_canonical<double, 53> = two 32-bit draws; Marsaglia polar method that
    returns y*mult first and keeps x*mult for the next call).  deal.II's SolverIDR fills its shadow vectors Q_1..Q_{s-1}
    with this stream, entry by entry in DoF order, continuing across the vectors."""
    rs = np.random.RandomState(5489)                      # init_genrand(5489) = std::mt19937's default seeding
    out = np.empty(n)
    k = 0
    buf = np.empty(0, dtype=np.uint64); pos = 0

    def canonical():
        nonlocal buf, pos
        if pos + 2 > len(buf):
            buf = rs.randint(0, 2 ** 32, size=1 << 16, dtype=np.uint64); pos = 0
        lo, hi = int(buf[pos]), int(buf[pos + 1]); pos += 2
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        return r if r < 1.0 else np.nextafter(1.0, 0.0)
    while k < n:
        while True:
            xx = 2.0 * canonical() - 1.0
            yy = 2.0 * canonical() - 1.0
            r2 = xx * xx + yy * yy
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = np.sqrt(-2.0 * np.log(r2) / r2)
        out[k] = yy * mult; k += 1
        if k < n:
            out[k] = xx * mult; k += 1
    return out


def idr(op, prec, b, x0, tol, s, max_its, normals=None, history=None):
    """deal.II 9.6 SolverIDR(AdditionalData(s)) with (right) preconditioning (local_smoothing.cc:1286-1297), recalled
    from the public source: shadow vectors Q_0 = r_0, Q_i = orthonormalised normal random vectors;
    per outer step s inner steps (one preconditioner + one operator application each) with an early exit as soon as
    the residual norm passes the tolerance, then one more application for omega.
    Returns (x, iterations, last_residual, converged)."""
    n = len(b)
    x = x0.copy()
    r = b - op(x)
    res = np.linalg.norm(r)
    if history is not None:
        history.append(res)
    step = 0
    if res <= tol:
        return x, 0, res, True
    if normals is None:
        normals = libstdcxx_normal_stream((s - 1) * n)
    G = [np.zeros(n) for _ in range(s)]
    U = [np.zeros(n) for _ in range(s)]
    Q = []
    for i in range(s):
        q = r.copy() if i == 0 else normals[(i - 1) * n:i * n].copy()
        for j in range(i):
            v = Q[j] * ((Q[j] @ q) / (q @ q))
            q = q - v
        if i != 0:
            q = q / np.linalg.norm(q)
        Q.append(q)
    M = np.eye(s)
    omega = 1.0
    while True:
        step += 1
        phi = np.array([Q[i] @ r for i in range(s)])
        for k in range(s):
            gamma = np.linalg.solve(M[k:, k:], phi[k:])
            v = r.copy()
            for j, i in enumerate(range(k, s)):
                v -= gamma[j] * G[i]
            uhat = prec(v) * omega
            for j, i in enumerate(range(k, s)):
                uhat += gamma[j] * U[i]
            ghat = op(uhat)
            if k > 0:
                alpha = (Q[0] @ ghat) / M[0, 0]
                for i in range(1, k):
                    ghat -= alpha * G[i - 1]
                    uhat -= alpha * U[i - 1]
                    alpha = (Q[i] @ ghat) / M[i, i]
                ghat -= alpha * G[k - 1]
                uhat -= alpha * U[k - 1]
            G[k] = ghat
            M[k, k] = G[k] @ Q[k]
            beta = phi[k] / M[k, k]
            r = r - beta * G[k]
            x = x + beta * uhat
            U[k] = uhat
            res = np.linalg.norm(r)
            if history is not None:
                history.append(res)
            if res <= tol:
                return x, step, res, True
            if step >= max_its:
                return x, step, res, False
            if k + 1 < s:
                phi[:k + 1] = 0.0
                for i in range(k + 1, s):
                    M[i, k] = Q[i] @ G[k]
                    phi[i] -= beta * M[i, k]
        uhat = prec(r)
        v = op(uhat)
        omega = (v @ r) / (v @ v)
        r = r - omega * v
        res = np.linalg.norm(r)
        x = x + omega * uhat
        if history is not None:
            history.append(res)
        if res <= tol:
            return x, step, res, True
        if step >= max_its:
            return x, step, res, False


# ---------------------------------------------------------------------------
class BlockGMGStokes:
    """The cheap-phase solver of StokesMatrixFreeHandlerLocalSmoothingImplementation::solve()."""

    def __init__(self, backend):
        be = self.be = backend
        self.mg_A = VCycle(be.n_levels, be.apply_A, be.dinv_A, be.prolongate_u, be.restrict_u, 1)
        self.mg_S = VCycle(be.n_levels, be.apply_M, be.dinv_M, be.prolongate_p, be.restrict_p, 4)
        self.n_u, self.n_p = be.n_u, be.n_p
        self.n_prec = 0

    def op(self, x):
        yu, yp = self.be.stokes(x[:self.n_u], x[self.n_u:])
        return np.concatenate([yu, yp])

    def prec(self, r):
        """BlockSchurPreconditioner::vmult (block_stokes_preconditioner.h:220-255)."""
        self.n_prec += 1
        zp = -self.mg_S.vmult(r[self.n_u:])
        t = r[:self.n_u] - self.be.BT(zp)
        zu = self.mg_A.vmult(t)
        return np.concatenate([zu, zp])

    def prec_expensive(self, r, a_tol=1e-2, s_tol=1e-6):
        """BlockSchurPreconditioner with do_solve_A = do_solve_Schur = true: inner CG solves preconditioned by the
        V-cycles (block_stokes_preconditioner.h:108-135: 10000 its, |src| * A-block tolerance; :312-345: 100 its,
        |src| * S-block tolerance, skipped when |src| <= 1e-50)."""
        rp = r[self.n_u:]
        zp = np.zeros(self.n_p)
        if np.linalg.norm(rp) > 1e-50:
            zp, its, ok = cg(self.be.M_active, self.mg_S.vmult, rp, np.linalg.norm(rp) * s_tol, 100)
            self.n_S_cg += its
            if not ok:
                raise RuntimeError("iterative (bottom right) solver did not converge")
        zp = -zp
        t = r[:self.n_u] - self.be.BT(zp)
        if getattr(self, "A_block_is_symmetric", True):
            zu, its, ok = cg(self.be.A_active, self.mg_A.vmult, t, np.linalg.norm(t) * a_tol, 10000)
        else:
            zu, its, ok = bicgstab(self.be.A_active, self.mg_A.vmult, t, np.linalg.norm(t) * a_tol, 10000)
        self.n_A_cg += its
        if not ok:
            raise RuntimeError("iterative (top left) solver did not converge")
        return np.concatenate([zu, zp])

    def solve(self, rhs, x0=None, tol_rel=1e-6, restart=100, max_its=200, history=None, krylov="gmres", idr_s=2,
              n_expensive=0, a_tol=1e-2, s_tol=1e-6):
        """Cheap phase (GMRES or IDR(s), at most max_its steps; 0 skips it, local_smoothing.cc:1270-1272), then on
        failure the expensive phase: FGMRES(restart) with inner CG solves, at most n_expensive steps (:1317-1350)."""
        if x0 is None:
            x0 = np.zeros_like(rhs)
        # tolerance: local_smoothing.cc:1016-1025 (velocity part of the guess zeroed)
        g = x0.copy(); g[:self.n_u] = 0.0
        r = self.op(g)
        ru = np.linalg.norm(rhs[:self.n_u] - r[:self.n_u])
        rp = np.linalg.norm(rhs[self.n_u:])
        tol = tol_rel * np.sqrt(ru * ru + rp * rp)
        self.n_prec = 0
        self.n_A_cg = self.n_S_cg = 0
        its, its_exp, ok, x, res = 0, 0, False, x0, np.nan
        if max_its > 0:
            if krylov == "gmres":
                x, its, res, ok = gmres_right(self.op, self.prec, rhs, x0, tol, restart, max_its, history)
            else:
                x, its, res, ok = idr(self.op, self.prec, rhs, x0, tol, idr_s, max_its, history=history)
        if not ok and n_expensive > 0:
            x, its_exp, res, ok = fgmres(self.op, lambda v: self.prec_expensive(v, a_tol, s_tol), rhs, x, tol, restart, n_expensive, history)
        return dict(x=x, iterations=its, iterations_expensive=its_exp, residual=res, converged=ok, tolerance=tol, n_prec=self.n_prec,
                    n_A_cg=self.n_A_cg, n_S_cg=self.n_S_cg)
