"""Device time of one prec_A / prec_Schur V-cycle and of the single operators at refinement R (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic

for R in [int(a) for a in sys.argv[1:]] or [4, 5, 6]:
    hier = synthetic.BoxHierarchy(3, R)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    h.build_preconditioner()
    h.setup_smoothers()
    src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
    rng = np.random.default_rng(0)
    xu = rng.uniform(-1, 1, hier.active.n_u); xu[hier.active.constrained_u("zero").astype(np.int64)] = 0
    src[0].from_host(xu); src[1].from_host(rng.uniform(-1, 1, hier.active.n_p))

    def timed(f, reps=20):
        import time
        for _ in range(3):
            f()
        h.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        h.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {"R": R,
           "vcycle_A_ms": timed(lambda: h.prec_A.vmult(dst[0], src[0])),
           "vcycle_S_ms": timed(lambda: h.prec_Schur.vmult(dst[1], src[1])),
           "A_ms": timed(lambda: h.A_block_matrix.vmult(dst[0], src[0])),
           "mass_ms": timed(lambda: h.Schur_complement_block_matrix.vmult(dst[1], src[1])),
           "stokes_ms": timed(lambda: h.stokes_matrix.vmult(dst, src)),
           "BT_ms": timed(lambda: h.BT_block.vmult(dst, src))}
    print(out, flush=True)
    h.close()
