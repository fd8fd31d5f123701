"""Short command for ncu captures: set up nsinker at refinement R and apply each operator a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
avg = sys.argv[3] if len(sys.argv) > 3 else "harmonic average"
hier = synthetic.BoxHierarchy(3, R)
eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
h = StokesMatrixFreeHandler(3)
h.setup_dofs(hier, "zero")
h.assemble(eta_q, avg, ps)
h.build_preconditioner()
src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
rng = np.random.default_rng(0)
xu = rng.uniform(-1, 1, hier.active.n_u); xu[hier.active.constrained_u("zero").astype(np.int64)] = 0
src[0].from_host(xu); src[1].from_host(rng.uniform(-1, 1, hier.active.n_p))
for _ in range(reps):
    h.A_block_matrix.vmult(dst[0], src[0])
    h.stokes_matrix.vmult(dst, src)
    h.BT_block.vmult(dst, src)
    h.Schur_complement_block_matrix.vmult(dst[1], src[1])
h.synchronize()
print("done")
