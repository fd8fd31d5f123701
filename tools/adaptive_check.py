#!/usr/bin/env python
"""Lab check of the adaptive path at bench size (GPU): convection-box shape, 3 global + 3 adaptive levels."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aspect_b200 import StokesMatrixFreeHandler, adaptive as padapt, benchmarks

base, rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3, int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.time()
hier = padapt.AdaptiveBoxHierarchy.refined(3, base, lambda x, l: min(x[2], 1.0 - x[2]) < 0.25 / (1 << (l - base)), rounds)
print(f"mesh {time.time() - t0:.1f} s: {hier.active.n_cells} cells, {hier.active.n_u}+{hier.active.n_p} DoFs, levels {[L.n_cells for L in hier.levels]}, "
      f"{len(hier.active.hanging_u)} hanging velocity nodes", flush=True)
eta, rhs, ps = benchmarks.convection_box_problem(hier)
h = StokesMatrixFreeHandler(3)
t0 = time.time()
h.setup_dofs(hier, "tangential")
h.assemble(eta, "none", ps)
h.build_preconditioner()
print(f"device set-up {time.time() - t0:.1f} s", flush=True)
h.parameters.update({"Linear solver tolerance": 1e-7, "Number of cheap Stokes solver steps": 200, "Maximum number of expensive Stokes solver steps": 0})
h.condense_rhs(rhs)
rhs[hier.active.constrained_u("tangential").astype(np.int64)] = 0.0
for k in range(3):
    t0 = time.time()
    sol, out = h.solve(rhs)
    print(f"solve {k}: {out['iterations']} iterations, converged {out['converged']}, {time.time() - t0:.3f} s, |u| = {np.linalg.norm(sol[:hier.active.n_u]):.6e}", flush=True)
# uniform mesh of the same finest resolution for comparison of the iteration count
from aspect_b200 import synthetic
if base + rounds <= 6:
    hu = synthetic.BoxHierarchy(3, base + rounds)
    eta_u, rhs_u, ps_u = benchmarks.convection_box_problem(hu)
    g = StokesMatrixFreeHandler(3)
    g.setup_dofs(hu, "tangential"); g.assemble(eta_u, "none", ps_u); g.build_preconditioner()
    g.parameters.update(h.parameters)
    t0 = time.time(); _, ou = g.solve(rhs_u); t0 = time.time(); _, ou = g.solve(rhs_u)
    print(f"uniform refinement {base + rounds}: {hu.active.n_u + hu.active.n_p} DoFs, {ou['iterations']} iterations, {time.time() - t0:.3f} s")
