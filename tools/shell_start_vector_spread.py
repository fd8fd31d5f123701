"""How much of the spherical-shell iteration count hangs on the DoF NUMBERING behind the Chebyshev start vector?

tests/spherical_nsinker_gmg_gc (106 iterations) is a 4-rank run of the global-coarsening handler: its level DoFs are the active
DoFs of separately partitioned coarse triangulations, numbered by deal.II in an order this repository cannot reconstruct.  The
numbering enters the algorithm in exactly one place -- PreconditionChebyshev's eigenvalue estimate starts from `global index
mod 11` (App. A.2) -- so this script repeats the oracle's solve of that golden with the start vector taken from random
renumberings (one permutation per level and block, seeded) and prints the iteration counts.  CPU only (oracle/); ~45 s per run.

    python tools/shell_start_vector_spread.py [n_trials]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import assembled, coeff, shell as oshell, solver  # noqa: E402


def solve(o, seed):
    L = o.active
    eo, ro = coeff.nsinker_shell(L.xq)
    active, levels, _, _ = coeff.viscosity_tables(o, eo, "harmonic")
    xm, det = L.midpoint_geometry()
    ps = coeff.pressure_scaling_shell(L, coeff.nsinker_shell(xm)[0], det)
    g = -L.xq / np.linalg.norm(L.xq, axis=-1, keepdims=True)
    f = assembled.rhs_force(L, coeff.average(ro, "harmonic")[..., None] * g)
    f[L.constrained_u("zero")] = 0.0
    be = solver.AssembledBackend(o, active, levels, ps, "zero")
    rng = np.random.default_rng(seed) if seed is not None else None
    orig = solver.Chebyshev.estimate_eigenvalues
    lam = []

    def patched(self):
        if rng is None:
            orig(self)
        else:
            n = len(self.dinv)
            ar = np.arange
            try:
                np_arange = solver.np.arange
                perm = rng.permutation(n)
                solver.np.arange = lambda m, *a, **k: perm if (m == n and not a and not k) else np_arange(m, *a, **k)
                orig(self)
            finally:
                solver.np.arange = np_arange
        lam.append(self.max_eig_est)
    solver.Chebyshev.estimate_eigenvalues = patched
    try:
        out = solver.BlockGMGStokes(be).solve(np.concatenate([f, np.zeros(L.n_p)]), tol_rel=1e-6, restart=200, max_its=2000)
    finally:
        solver.Chebyshev.estimate_eigenvalues = orig
    return out["iterations"], lam


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    o = oshell.ShellHierarchy(1)
    its0, lam0 = solve(o, None)
    print(f"numbering of this repository: {its0} iterations; lambda_max estimates {['%.4f' % x for x in lam0]}", flush=True)
    counts = []
    for s in range(n_trials):
        its, lam = solve(o, 1000 + s)
        counts.append(its)
        print(f"random renumbering {s}: {its} iterations; lambda_max estimates {['%.4f' % x for x in lam]}", flush=True)
    print(f"spread over {n_trials} renumberings: {min(counts)} .. {max(counts)} (reference golden, 4 ranks, GMG-GC: 106; this numbering: {its0})")
