#!/usr/bin/env python
"""Generates tests/golden/nsinker_R{R}_cpu_reference.npz: one COMPLETE block-GMG solve of the bench's headline configuration
(nsinker, 8 sinkers, DR 1e4, harmonic average, no-slip box, GMRES(100), tol 1e-6) by the C++ host-core implementation of the
reference algorithm (oracle/csrc/smf_cpu_solver.cc), reduced to what a GPU test can compare at that size without the oracle
having to repeat the solve on the GPU box: iteration counts, tolerance, residuals, norms and seeded samples of the solution.

    python tools/make_full_size_fixture.py 7        # ~50 GB of host memory, several minutes on 8 cores
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coeff, cpuref                      # noqa: E402
from aspect_b200 import benchmarks, synthetic         # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 7
t0 = time.time()
hier = synthetic.BoxHierarchy(3, R)
eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, "harmonic")
del eta_q
cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
threads = cpuref.set_num_threads()
print(f"set-up {time.time() - t0:.1f} s, {threads} threads, SIMD width {cpuref.lib().smfc_width()}", flush=True)
out = cb.solve(rhs, tol_rel=1e-6, restart=100, max_its=200)
print({k: v for k, v in out.items() if k != "x"}, flush=True)
nu, npr = hier.active.n_u, hier.active.n_p
u, p = out["x"][:nu], out["x"][nu:] * ps
p = p - p.mean()                                      # the pressure is defined up to a constant on the no-slip box
rng = np.random.default_rng(2026)
iu, ip = np.sort(rng.choice(nu, 4096, replace=False)), np.sort(rng.choice(npr, 4096, replace=False))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"nsinker_R{R}_cpu_reference.npz"),
                    refinement=R, n_u=nu, n_p=npr, iterations=out["iterations"], n_prec=out["n_prec"], tolerance=out["tolerance"],
                    initial_residual=out["initial_residual"], final_residual=out["residual"], viscosity_range=np.array([emin, emax]),
                    u_norm=np.linalg.norm(u), p_norm=np.linalg.norm(p), rhs_norm=np.linalg.norm(rhs),
                    u_index=iu, u_sample=u[iu], p_index=ip, p_sample=p[ip], seconds=out["seconds"], threads=threads)
print("written", flush=True)
