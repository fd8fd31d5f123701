"""Experiment: A block at refinement R through the production brick kernel and through the TMA-staged lattice kernel
(SMF_OPT_LATTICE_VELOCITY); kernel time alone from the device-side profiler (CUDA events around the launch).
    python tools/lab_lattice_velocity.py R [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hier = synthetic.BoxHierarchy(3, R)
eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
x = np.random.default_rng(0).uniform(-1, 1, hier.active.n_u)
x[hier.active.constrained_u("zero").astype(np.int64)] = 0
ref = None
for lattice in (False, True):
    h = StokesMatrixFreeHandler(3)
    h.use_lattice_velocity = lattice
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    h.build_preconditioner()
    src, dst = h.initialize_dof_vector(0), h.initialize_dof_vector(0)
    src.from_host(x)
    for _ in range(3):
        h.A_block_matrix.vmult(dst, src)
    h.profile_enable(1)
    for _ in range(reps):
        h.A_block_matrix.vmult(dst, src)
    ms, cnt = h.profile_read()["A_fine"]
    y = dst.to_host()
    if ref is None:
        ref = y
    n_u, n_cells = hier.active.n_u, hier.active.n_cells
    print({"R": R, "lattice_kernel": lattice, "A_kernel_us": 1e3 * ms / max(cnt, 1), "launches": cnt,
           "GDoF/s": n_u / (ms / max(cnt, 1) * 1e-3) / 1e9, "algorithmic_GB/s": (16 * n_u + 8 * n_cells) / (ms / max(cnt, 1) * 1e-3) / 1e9,
           "rel_diff_vs_brick_kernel": float(np.linalg.norm(y - ref) / np.linalg.norm(ref))}, flush=True)
    h.close()
