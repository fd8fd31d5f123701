"""CPU oracle for the matrix-free GMG Stokes path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the algorithm of ASPECT's
`Stokes solver type = block GMG` path
(/root/reference/source/simulator/solver/matrix_free_operators.cc,
 stokes_matrix_free_local_smoothing.cc, block_stokes_preconditioner.h)
together with the deal.II 9.6 semantics those files rely on (deal.II is an
external dependency that is NOT vendored in the reference tree and cannot be
built here; see DESIGN.md "Oracle").

It is the checker, never the product:
  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
    `--impl reference` legs of `bench.py` may import it;
  * nothing under `aspect_b200/` imports it, and the product fails loudly when
    its CUDA library is missing.

Pinning (see tests/test_oracle_goldens.py): the oracle reproduces the golden
outputs the reference's own integration tests hold for this path
(DoF counts, viscosity ranges, outer iteration counts of
tests/sol_cx_{2,4}_gmg, tests/nsinker_gmg_gc, tests/nsinker_gmg_no_avg and the
SolCx analytic errors to all printed digits).  Operator-level parity (a single
vmult) is unpinned by the reference (it has no operator-level known-answer
test); it is pinned here through an independently assembled sparse matrix.
"""
