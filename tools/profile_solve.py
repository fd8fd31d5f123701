"""Short command for ncu launch lists: one block-GMG solve of nsinker at refinement R (no warm-up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

R = int(sys.argv[1]) if len(sys.argv) > 1 else 6
hier, h, rhs, _, _ = bench.build_problem(R, 0)
n_u = hier.active.n_u
rhs_dev, sol_dev = h.initialize_dof_vector(), h.initialize_dof_vector()
rhs_dev[0].from_host(rhs[:n_u]); rhs_dev[1].from_host(rhs[n_u:])
out = h.solve_device(sol_dev, rhs_dev)
print(out)
