"""Full-size checks (runs last): the operators, transfers and V-cycles of the refinement-7 nsinker-shaped problem
(2,097,152 cells, 53 M DoFs -- the size bench.py times) through size-independent properties, since no oracle run fits:
symmetry and linearity of A / Stokes / pressure mass, the closed form 1^T M 1 = sum_cells ps^2 vol / eta_cell, adjointness of
restriction and prolongation, prolongation of a constant pressure, and the symmetry of the two V-cycle preconditioners
(equal Chebyshev polynomials before and after the coarse correction make PreconditionMG a symmetric operator; the oracle has
this property to 1e-14 at small sizes).  A wrong kernel shows up as an O(1) violation; the tolerances only absorb the
round-off of 5e7-term sums.  SMF_FULL_SIZE_R overrides the refinement."""
import os

import numpy as np
import pytest

from aspect_b200 import SMF_ACTIVE, StokesMatrixFreeHandler, synthetic

pytestmark = pytest.mark.gpu
R = int(os.environ.get("SMF_FULL_SIZE_R", "7"))


def _dot_err(a, b, scale):
    return abs(a - b) / scale


def test_full_size_properties():
    hier = synthetic.BoxHierarchy(3, R)
    act = hier.active
    rng = np.random.default_rng(2026)
    eta_cell = 10.0 ** rng.uniform(-2.0, 2.0, size=(act.n_cells, 1))            # four orders of magnitude, cell to cell
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "zero")
    ps = 3.0
    h.assemble(np.repeat(eta_cell, 27, axis=1), "harmonic average", ps)
    h.build_preconditioner()
    h.setup_smoothers()
    cons = act.constrained_u("zero").astype(np.int64)
    nu, npr = act.n_u, act.n_p

    def rand_u():
        v = rng.uniform(-1.0, 1.0, nu); v[cons] = 0.0
        return v
    x, y, xp, yp = rand_u(), rand_u(), rng.uniform(-1.0, 1.0, npr), rng.uniform(-1.0, 1.0, npr)
    sx, sy, d = h.initialize_dof_vector(), h.initialize_dof_vector(), h.initialize_dof_vector()
    sx[0].from_host(x); sx[1].from_host(xp); sy[0].from_host(y); sy[1].from_host(yp)

    # A block: symmetry and linearity
    h.A_block_matrix.vmult(d[0], sx[0]); Ax = d[0].to_host()
    h.A_block_matrix.vmult(d[0], sy[0]); Ay = d[0].to_host()
    assert _dot_err(y @ Ax, x @ Ay, np.linalg.norm(y) * np.linalg.norm(Ax)) < 1e-11
    comb = h.initialize_dof_vector(0).from_host(2.0 * x - 3.0 * y)
    h.A_block_matrix.vmult(d[0], comb)
    assert np.linalg.norm(d[0].to_host() - (2.0 * Ax - 3.0 * Ay)) < 1e-12 * (np.linalg.norm(Ax) + np.linalg.norm(Ay))
    assert x @ Ax > 0.0
    comb.free()

    # Stokes operator: symmetric as a block operator
    h.stokes_matrix.vmult(d, sx); Kx = (d[0].to_host(), d[1].to_host())
    h.stokes_matrix.vmult(d, sy); Ky = (d[0].to_host(), d[1].to_host())
    a, b = y @ Kx[0] + yp @ Kx[1], x @ Ky[0] + xp @ Ky[1]
    scale = np.sqrt(y @ y + yp @ yp) * np.sqrt(Kx[0] @ Kx[0] + Kx[1] @ Kx[1])
    assert _dot_err(a, b, scale) < 1e-11
    # ... and B^T is the transpose of the pressure rows: <y, B^T xp> = <B y, xp> with B y = pressure rows of Stokes(y, 0)
    h.BT_block.vmult(d, sx); BTxp = d[0].to_host()
    zero_p = h.initialize_dof_vector(1).zero()
    h.stokes_matrix.vmult(d, (sy[0], zero_p)); By = d[1].to_host()
    assert _dot_err(y @ BTxp, By @ xp, np.linalg.norm(y) * np.linalg.norm(BTxp)) < 1e-11

    # pressure mass matrix: symmetry and the closed form of 1^T M 1
    h.Schur_complement_block_matrix.vmult(d[1], sx[1]); Mx = d[1].to_host()
    h.Schur_complement_block_matrix.vmult(d[1], sy[1]); My = d[1].to_host()
    assert _dot_err(yp @ Mx, xp @ My, np.linalg.norm(yp) * np.linalg.norm(Mx)) < 1e-11
    one = h.initialize_dof_vector(1).from_host(np.ones(npr))
    h.Schur_complement_block_matrix.vmult(d[1], one)
    want = float((ps * ps / eta_cell[:, 0]).sum() * np.prod(act.h))
    assert abs(d[1].to_host().sum() / want - 1.0) < 1e-10

    # transfers between the two finest levels: adjointness, constants
    L, C = R, hier.levels[R - 1]
    cu, cp = rng.uniform(-1.0, 1.0, C.n_u), rng.uniform(-1.0, 1.0, C.n_p)
    cu[C.constrained_u("zero").astype(np.int64)] = 0.0
    fu, fp = h.initialize_dof_vector(0, L), h.initialize_dof_vector(1, L)
    h.prolongate(0, L, fu, h.initialize_dof_vector(0, L - 1).from_host(cu)); Pcu = fu.to_host()
    h.prolongate(1, L, fp, h.initialize_dof_vector(1, L - 1).from_host(cp)); Pcp = fp.to_host()
    ru, rp = h.initialize_dof_vector(0, L - 1).zero(), h.initialize_dof_vector(1, L - 1).zero()
    h.restrict_and_add(0, L, ru, sy[0]); h.restrict_and_add(1, L, rp, sy[1])
    assert _dot_err(y @ Pcu, cu @ ru.to_host(), np.linalg.norm(y) * np.linalg.norm(Pcu)) < 1e-11
    assert _dot_err(yp @ Pcp, cp @ rp.to_host(), np.linalg.norm(yp) * np.linalg.norm(Pcp)) < 1e-11
    h.prolongate(1, L, fp, h.initialize_dof_vector(1, L - 1).from_host(np.ones(C.n_p)))
    assert np.abs(fp.to_host() - 1.0).max() < 1e-14

    # the two V-cycle preconditioners are symmetric operators
    h.prec_A.vmult(d[0], sx[0]); Gx = d[0].to_host()
    h.prec_A.vmult(d[0], sy[0]); Gy = d[0].to_host()
    assert _dot_err(y @ Gx, x @ Gy, np.linalg.norm(y) * np.linalg.norm(Gx)) < 1e-9
    assert x @ Gx > 0.0
    h.prec_Schur.vmult(d[1], sx[1]); Sx = d[1].to_host()
    h.prec_Schur.vmult(d[1], sy[1]); Sy = d[1].to_host()
    assert _dot_err(yp @ Sx, xp @ Sy, np.linalg.norm(yp) * np.linalg.norm(Sx)) < 1e-9
    h.close()
