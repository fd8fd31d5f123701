"""BASELINE cfg 5 refines ONE coarse cell R + 1 times; the multi-GPU runs of this repository refine 2 x 2 x 2 coarse cells R times
(one tree per rank).  Same active mesh, one multigrid level fewer.  This script solves both set-ups of the same nsinker problem on
ONE GPU and prints the iteration counts:   python tools/coarse_grid_equivalence.py [R]   (R + 1 = refinement of the single cell)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic  # noqa: E402


def solve(refinement, repetitions):
    hier = synthetic.BoxHierarchy(3, refinement, None, repetitions, 0, 1)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-6, "Number of cheap Stokes solver steps": 200, "GMRES solver restart length": 100})
    _, out = h.solve(rhs)
    n = hier.active.n_u + hier.active.n_p
    h.close()
    return n, len(hier.levels), out["iterations"], out["converged"]


if __name__ == "__main__":
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    a = solve(R + 1, (1, 1, 1))
    b = solve(R, (2, 2, 2))
    print(f"1 coarse cell refined {R + 1} times: {a[0]} DoFs, {a[1]} levels, {a[2]} iterations (converged {a[3]}); "
          f"2x2x2 coarse cells refined {R} times: {b[0]} DoFs, {b[1]} levels, {b[2]} iterations (converged {b[3]})")
