"""GPU parity: every operator of the C ABI against the oracle on the same seeded inputs.

Reference functions under test (replaced by CUDA kernels):
  StokesOperator::vmult, ABlockOperator::vmult, BTBlockOperator::vmult, MassMatrixOperator::vmult,
  compute_diagonal (matrix_free_operators.cc), MGTransfer prolongate / restrict_and_add.
Tolerance: relative 1e-12 in double (BASELINE.json north_star); measured errors are ~1e-15.
"""
import numpy as np
import pytest

from oracle import mesh as omesh
from aspect_b200 import StokesMatrixFreeHandler, SMF_ACTIVE, synthetic
from util import oracle_backend, rel, seeded

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [(2, 1, "tangential"), (2, 3, "tangential"), (2, 4, "zero"), (3, 0, "zero"), (3, 1, "zero"), (3, 2, "tangential"), (3, 3, "zero")]


def _setup(dim, R, bc, averaging, ps=3.7):
    hier = synthetic.BoxHierarchy(dim, R)
    ohier = omesh.Hierarchy(dim, R)
    rng = np.random.default_rng(42 + dim * 10 + R)
    eta_q = np.exp(rng.uniform(-4, 4, size=(hier.active.n_cells, 3 ** dim)))
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, averaging, ps)
    h.build_preconditioner()
    ob, *_ = oracle_backend(ohier, eta_q, averaging, ps, bc)
    return hier, h, ob


@pytest.mark.parametrize("dim,R,bc", CASES)
@pytest.mark.parametrize("averaging", ["none", "harmonic average"])
def test_active_operators(dim, R, bc, averaging):
    hier, h, ob = _setup(dim, R, bc, averaging)
    act = hier.active
    xu, xp = seeded(act.n_u, 1), seeded(act.n_p, 2)
    src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
    src[0].from_host(xu); src[1].from_host(xp)
    h.stokes_matrix.vmult(dst, src)
    yu, yp = ob.stokes(xu, xp)
    assert rel(dst[0].to_host(), yu) < TOL and rel(dst[1].to_host(), yp) < TOL
    h.BT_block.vmult(dst, src)
    assert rel(dst[0].to_host(), ob.BT(xp)) < TOL
    h.A_block_matrix.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), ob.A_active(xu)) < TOL
    h.Schur_complement_block_matrix.vmult(dst[1], src[1])
    assert rel(dst[1].to_host(), ob.M_active(xp)) < TOL
    h.close()


@pytest.mark.parametrize("dim,R,bc", CASES)
def test_level_operators_diagonals_transfers(dim, R, bc):
    hier, h, ob = _setup(dim, R, bc, "harmonic average")
    for l, L in enumerate(hier.levels):
        xu, xp = seeded(L.n_u, 10 + l), seeded(L.n_p, 20 + l)
        su, du = h.initialize_dof_vector(0, l).from_host(xu), h.initialize_dof_vector(0, l)
        sp, dp = h.initialize_dof_vector(1, l).from_host(xp), h.initialize_dof_vector(1, l)
        h.mg_matrices_A_block[l].vmult(du, su)
        assert rel(du.to_host(), ob.apply_A(l, xu)) < TOL
        h.mg_matrices_Schur_complement[l].vmult(dp, sp)
        assert rel(dp.to_host(), ob.apply_M(l, xp)) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 0), ob.dinv_A(l)) < TOL
        assert rel(h.get_matrix_diagonal_inverse(l, 1), ob.dinv_M(l)) < TOL
        if l > 0:
            C = hier.levels[l - 1]
            cu, cp = seeded(C.n_u, 30 + l), seeded(C.n_p, 40 + l)
            scu, scp = h.initialize_dof_vector(0, l - 1).from_host(cu), h.initialize_dof_vector(1, l - 1).from_host(cp)
            h.prolongate(0, l, du, scu)
            assert rel(du.to_host(), ob.prolongate_u(l, cu)) < TOL
            h.prolongate(1, l, dp, scp)
            assert rel(dp.to_host(), ob.prolongate_p(l, cp)) < TOL
            rcu = h.initialize_dof_vector(0, l - 1); rcp = h.initialize_dof_vector(1, l - 1)
            h.restrict_and_add(0, l, rcu, su)
            assert rel(rcu.to_host(), ob.restrict_u(l, xu)) < TOL
            h.restrict_and_add(1, l, rcp, sp)
            assert rel(rcp.to_host(), ob.restrict_p(l, xp)) < TOL
    h.close()


@pytest.mark.parametrize("averaging", ["none", "harmonic average"])
@pytest.mark.parametrize("grid_cap,pipelined", [(0, True), (3, True), (1, True), (0, False)])
def test_pipelined_velocity_kernel(averaging, grid_cap, pipelined):
    """Structured levels run the persistent, cp.async-pipelined kernel; a capped grid makes every block walk several
    bricks (16 bricks at R = 3: 6 or 16 per block) so that the prefetch / buffer rotation is exercised.  Same 1e-12
    parity against the oracle as the one-brick-per-block kernel (pipelined = False)."""
    dim, R, bc = 3, 3, "zero"
    hier = synthetic.BoxHierarchy(dim, R)
    ohier = omesh.Hierarchy(dim, R)
    eta_q = np.exp(np.random.default_rng(7).uniform(-4, 4, size=(hier.active.n_cells, 27)))
    h = StokesMatrixFreeHandler(dim)
    h.use_pipelined_kernel = pipelined
    h.max_grid_blocks = grid_cap
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, averaging, 2.5)
    h.build_preconditioner()
    ob, *_ = oracle_backend(ohier, eta_q, averaging, 2.5, bc)
    act = hier.active
    xu, xp = seeded(act.n_u, 3), seeded(act.n_p, 4)
    src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
    src[0].from_host(xu); src[1].from_host(xp)
    for _ in range(2):      # twice: the result must not depend on what the destination held before
        h.A_block_matrix.vmult(dst[0], src[0])
        assert rel(dst[0].to_host(), ob.A_active(xu)) < TOL
    h.stokes_matrix.vmult(dst, src)
    yu, yp = ob.stokes(xu, xp)
    assert rel(dst[0].to_host(), yu) < TOL and rel(dst[1].to_host(), yp) < TOL
    for l in (2, 3):
        L = hier.levels[l]
        xl = seeded(L.n_u, 50 + l)
        su, du = h.initialize_dof_vector(0, l).from_host(xl), h.initialize_dof_vector(0, l)
        h.mg_matrices_A_block[l].vmult(du, su)
        assert rel(du.to_host(), ob.apply_A(l, xl)) < TOL
    h.close()


@pytest.mark.parametrize("dim,R,bc,averaging", [(2, 3, "tangential", "none"), (3, 2, "zero", "harmonic average"), (3, 3, "zero", "none")])
def test_correct_stokes_rhs(dim, R, bc, averaging):
    """StokesMatrixFreeHandler::correct_stokes_rhs (local_smoothing.cc:617-754): rhs -= Stokes(u0) with plain reads of
    u0 (boundary values kept) and nothing added to constrained rows, against the UNCONSTRAINED assembled operator."""
    from oracle import assembled, coeff
    hier = synthetic.BoxHierarchy(dim, R)
    ohier = omesh.Hierarchy(dim, R)
    rng = np.random.default_rng(11)
    eta_q = np.exp(rng.uniform(-3, 3, size=(hier.active.n_cells, 3 ** dim)))
    ps = 1.7
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, averaging, ps)
    h.build_preconditioner()
    act = ohier.active
    op = {"none": "none", "harmonic average": "harmonic"}[averaging]
    eta_active, _, _, _ = coeff.viscosity_tables(ohier, eta_q, op)
    K, _, _ = assembled.stokes_matrix(act, eta_active, ps, None)
    cons = hier.active.constrained_u(bc).astype(np.int64)
    u0 = np.zeros(act.n_u + act.n_p)
    u0[cons] = rng.uniform(-1, 1, len(cons))                       # inhomogeneous boundary values, zero elsewhere
    rhs = rng.uniform(-1, 1, act.n_u + act.n_p)
    corr = -(K @ u0)
    corr[cons] = 0.0
    want = rhs + corr
    r, u = h.initialize_dof_vector(), h.initialize_dof_vector()
    r[0].from_host(rhs[:act.n_u]); r[1].from_host(rhs[act.n_u:])
    u[0].from_host(u0[:act.n_u]); u[1].from_host(u0[act.n_u:])
    h.correct_stokes_rhs(r, u)
    got = np.concatenate([r[0].to_host(), r[1].to_host()])
    assert rel(got, want) < TOL
    assert np.array_equal(got[cons], rhs[cons])                     # constrained rows untouched
    h.close()
