"""Schur V-cycle alone at refinement R (for ncu / timing of the lattice path): python tools/profile_lattice.py R [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hier = synthetic.BoxHierarchy(3, R)
eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
h = StokesMatrixFreeHandler(3)
h.setup_dofs(hier, "zero")
h.assemble(eta_q, "harmonic average", ps)
h.build_preconditioner()
h.setup_smoothers()
src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
src[1].from_host(np.random.default_rng(0).uniform(-1, 1, hier.active.n_p))
for _ in range(3):
    h.prec_Schur.vmult(dst[1], src[1])
h.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    h.prec_Schur.vmult(dst[1], src[1])
h.synchronize()
print({"R": R, "lattice": h.lattice_schur_active(), "zpt": os.environ.get("SMF_LAT_ZPT", "auto"),
       "vcycle_S_ms": (time.perf_counter() - t0) / reps * 1e3}, flush=True)
h.close()
