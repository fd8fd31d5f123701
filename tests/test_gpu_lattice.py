"""Lattice path of the Schur-complement V-cycle (aspect_b200/csrc/schur_lattice.cu, SMF_OPT_LATTICE_SCHUR) against the
oracle's prec_Schur (PreconditionMG with Chebyshev(4) x 4 smoothing of the pressure mass matrix,
stokes_matrix_free_local_smoothing.cc:803-849, :893-916) and against the cell-batch path of the same library."""
import numpy as np
import pytest

from oracle import mesh as omesh, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import oracle_backend, rel, seeded

pytestmark = pytest.mark.gpu


def _handler(hier, eta_q, avg, ps, lattice, graphs=True):
    h = StokesMatrixFreeHandler(3)
    h.use_lattice_schur = lattice
    h.use_cuda_graphs = graphs
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, avg, ps)
    h.build_preconditioner()
    h.setup_smoothers()
    return h


@pytest.mark.parametrize("R,extents,reps", [(3, None, None), (4, None, None), (2, (2.0, 1.0, 0.5), (3, 2, 1))])
def test_lattice_vcycle_matches_oracle_and_batch_path(R, extents, reps):
    """prec_Schur on a seeded vector: lattice path = oracle (1e-10) = cell-batch path (1e-12); cubes and a 3 x 2 x 1-tree box
    with unequal cell sizes (nx != ny != nz)."""
    avg = "harmonic average"
    hier = synthetic.BoxHierarchy(3, R, extents, reps)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
    x = seeded(hier.active.n_p, 11)
    out = {}
    for lattice in (True, False):
        for graphs in (True, False):
            h = _handler(hier, eta_q, avg, ps, lattice, graphs)
            assert h.lattice_schur_active() == lattice
            src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
            src[1].from_host(x)
            h.prec_Schur.vmult(dst[1], src[1])
            h.prec_Schur.vmult(dst[1], src[1])          # (second call: graph replay)
            out[(lattice, graphs)] = dst[1].to_host()
            h.close()
    # replayed graph = direct launches; the two handlers estimate their smoother eigenvalues separately with the cell-batch
    # mass kernel (unordered REDs), so theta / delta differ in the last bits
    assert rel(out[(True, True)], out[(True, False)]) < 1e-13
    assert rel(out[(True, True)], out[(False, True)]) < 1e-12
    if reps is None:
        ohier = omesh.Hierarchy(3, R)
        ob, *_ = oracle_backend(ohier, eta_q, avg, ps, "zero")
        s = osolver.BlockGMGStokes(ob)
        assert rel(out[(True, True)], s.mg_S.vmult(x)) < 1e-10


def test_lattice_path_is_deterministic_and_solve_counts_match():
    """Two solves through the lattice path give bit-identical pressure V-cycles (fixed summation order, no REDs) and the same
    outer iteration count as the cell-batch path; per-point viscosities on the LEVELS (n_eta = 27) fall back to the cell-batch path."""
    avg = "harmonic average"
    hier = synthetic.BoxHierarchy(3, 3)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
    its = {}
    for lattice in (True, False):
        h = _handler(hier, eta_q, avg, ps, lattice)
        h.parameters["Linear solver tolerance"] = 1e-6
        sol, out = h.solve(rhs)
        its[lattice] = (out["iterations"], sol)
        if lattice:
            x = seeded(hier.active.n_p, 3)
            src, d1, d2 = h.initialize_dof_vector(), h.initialize_dof_vector(), h.initialize_dof_vector()
            src[1].from_host(x)
            h.prec_Schur.vmult(d1[1], src[1]); h.prec_Schur.vmult(d2[1], src[1])
            assert np.array_equal(d1[1].to_host(), d2[1].to_host())
        h.close()
    assert its[True][0] == its[False][0]
    nu = hier.active.n_u
    assert rel(its[True][1][:nu], its[False][1][:nu]) < 1e-8
    # `Material averaging = none`: 27 viscosities per ACTIVE cell, but the level tables the V-cycle sees are DGQ0 (one per cell,
    # local_smoothing.cc:211-294): the lattice path stays on; `project to Q1` puts 27 values on every level: cell-batch path
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "none")
    h = _handler(hier, eta_q, "none", ps, True)
    assert h.lattice_schur_active()
    h.close()
    h = _handler(hier, eta_q, "project to Q1", ps, True)
    assert not h.lattice_schur_active()
    h.close()
