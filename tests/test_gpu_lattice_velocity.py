"""SMF_OPT_LATTICE_VELOCITY (experimental): the A block of structured box levels through velocity_lat_kernel -- lattice-ordered
copies of the vectors, bricks staged by TMA box loads -- against the production brick kernel and the oracle
(ABlockOperator::local_apply, matrix_free_operators.cc:974-1052)."""
import numpy as np
import pytest

from oracle import mesh as omesh
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import oracle_backend, rel, seeded

pytestmark = pytest.mark.gpu


def _handler(hier, eta_q, avg, ps, bc, lattice, max_grid_blocks=0):
    h = StokesMatrixFreeHandler(3)
    h.use_lattice_velocity = lattice
    h.max_grid_blocks = max_grid_blocks
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, avg, ps)
    h.build_preconditioner()
    return h


@pytest.mark.parametrize("bc,grid", [("zero", 0), ("tangential", 0), ("none", 0), ("zero", 3)])
def test_lattice_A_block_matches_brick_kernel_and_oracle(bc, grid):
    """Active level and every structured level, three boundary-condition patterns (whole faces per component); grid = 3 forces
    several bricks per persistent block (the mbarrier ring is reused)."""
    avg, R = "harmonic average", 3
    hier = synthetic.BoxHierarchy(3, R)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
    ohier = omesh.Hierarchy(3, R)
    ob, *_ = oracle_backend(ohier, eta_q, avg, ps, bc)
    out = {}
    for lattice in (True, False):
        h = _handler(hier, eta_q, avg, ps, bc, lattice, grid)
        res = []
        for level in [-1] + list(range(2, R + 1)):
            L = hier.active if level < 0 else hier.levels[level]
            x = seeded(L.n_u, 7 + level)
            x[L.constrained_u(bc).astype(np.int64)] = 0.0
            src, dst = h.initialize_dof_vector(0, level), h.initialize_dof_vector(0, level)
            src.from_host(x)
            op = h.A_block_matrix if level < 0 else h.mg_matrices_A_block[level]
            op.vmult(dst, src)
            res.append(dst.to_host())
            if level < 0 and lattice:
                assert rel(res[-1], ob.A_active(x)) < 1e-12
        out[lattice] = res
        h.close()
    for a, b in zip(out[True], out[False]):
        assert rel(a, b) < 1e-13


def test_lattice_velocity_solve_matches():
    avg = "harmonic average"
    hier = synthetic.BoxHierarchy(3, 3)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, avg)
    sols = {}
    for lattice in (True, False):
        h = _handler(hier, eta_q, avg, ps, "zero", lattice)
        h.parameters["Linear solver tolerance"] = 1e-6
        sol, out = h.solve(rhs)
        sols[lattice] = (out["iterations"], sol)
        h.close()
    assert sols[True][0] == sols[False][0]
    nu = hier.active.n_u
    assert rel(sols[True][1][:nu], sols[False][1][:nu]) < 1e-8
