"""Adaptive meshes with hanging nodes, oracle side only (groundwork for SURVEY 8(a) a12; the CUDA path rejects them):
the restatement of deal.II's hanging-node constraints, MatrixFreeOperators::Base constraint convention and local
smoothing with refinement-edge / interface operators (oracle/adaptive.py) against what it must reduce to and satisfy."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import adaptive, assembled, coeff, mesh
from util import golden


@pytest.mark.parametrize("dim,R", [(2, 2), (3, 1)])
def test_uniform_mesh_reduces_to_the_uniform_restatement(dim, R):
    box = adaptive.AdaptiveBox.refined(dim, R, lambda x, l: False, 0)
    L = mesh.Level(dim, R)
    assert np.array_equal(box.cell_dofs_u, L.cell_dofs_u) and np.array_equal(box.cell_dofs_p, L.cell_dofs_p)
    assert not box.hanging_u and not box.hanging_p
    eta = np.random.default_rng(0).uniform(0.5, 2.0, (box.n_cells, 3 ** dim))
    ops = box.assemble(eta, 1.7, "zero")
    cons = L.constrained_u("zero")
    assert abs(ops["A"] - assembled.assemble_A(L, eta, cons)).max() < 1e-13
    assert abs(ops["BT"] - assembled.assemble_BT(L, 1.7, cons)).max() < 1e-14
    assert abs(ops["M"] - assembled.assemble_mass(L, eta, 1.7)).max() < 1e-14


@pytest.mark.parametrize("dim", [2, 3])
def test_hanging_node_constraints(dim):
    box = adaptive.AdaptiveBox.refined(dim, 1, lambda x, l: np.linalg.norm(x) < 0.6, 2)
    assert len({l for l, _ in box.cells}) >= 2 and box.hanging_u and box.hanging_p
    # 1D weights of SURVEY App. A.1 and their tensor products only
    w1 = {0.375, 0.75, -0.125, 1.0}
    allowed = {round(a * b, 10) for a in w1 for b in w1}
    assert {round(w, 10) for ws in box.hanging_u.values() for _, w in ws} <= allowed
    assert {round(w, 10) for ws in box.hanging_p.values() for _, w in ws} <= {0.5, 0.25}
    # a polynomial of degree 2 (1) per variable is continuous in the Q2 (Q1) space: it satisfies every constraint
    inv_u = {k: np.array(key) / (2 << box.Lmax) for key, k in box.node_u.items()}
    inv_p = {k: np.array(key) / (1 << box.Lmax) for key, k in box.node_p.items()}
    fu = lambda x: np.prod([1 + 0.3 * (d + 1) * x[d] + 0.7 * x[d] ** 2 for d in range(dim)])
    fp = lambda x: np.prod([1 + 0.3 * (d + 1) * x[d] for d in range(dim)])
    vu = np.array([fu(inv_u[k]) for k in range(len(inv_u))]); vp = np.array([fp(inv_p[k]) for k in range(len(inv_p))])
    assert max(abs(vu[k] - sum(w * vu[m] for m, w in ws)) for k, ws in box.hanging_u.items()) < 1e-14
    assert max(abs(vp[k] - sum(w * vp[m] for m, w in ws)) for k, ws in box.hanging_p.items()) < 1e-14
    # constrained operator: symmetric, rigid motions in the null space, divergence of u = x integrates to -dim
    ops = box.assemble(np.random.default_rng(1).uniform(0.5, 2.0, box.n_cells), 1.0, "none")
    A, cons = ops["A"], ops["constrained_u"]
    assert abs(A - A.T).max() < 1e-13
    X = np.array([inv_u[k] for k in range(len(inv_u))])
    rot = np.zeros(box.n_u); rot[0::dim] = X[:, 1]; rot[1::dim] = -X[:, 0]
    rot[cons] = 0.0                                   # constrained entries of a condensed vector are not read
    r = A @ rot; r[cons] = 0.0
    assert np.linalg.norm(r) < 1e-12
    u = np.zeros(box.n_u)
    for d in range(dim):
        u[d::dim] = X[:, d]
    u[cons] = 0.0
    assert abs((ops["BT"].T @ u).sum() + dim) < 1e-12


def test_local_smoothing_reduces_to_the_reference_golden():
    """No adaptive refinement: the adaptive V-cycle (copy indices, edge operators all empty) is the uniform one --
    tests/nsinker_gmg_gc, first solve: 21 iterations."""
    box = adaptive.AdaptiveBox.refined(3, 2, lambda x, l: False, 0)
    L = mesh.Level(3, 2)
    eta_q, rho_q = coeff.nsinker(box.quadrature_points(), 4, 1e4)
    f = assembled.rhs_gravity(L, rho_q, np.array([0.0, 0.0, -1.0]))
    f[L.constrained_u("zero")] = 0.0
    out = adaptive.block_gmg_solve(box, coeff.average(eta_q, "harmonic")[:, 0], np.concatenate([f, np.zeros(box.n_p)]), "zero", 1.0, 1e-6)
    assert out["converged"] and out["iterations"] == golden("nsinker_gmg_gc")["stokes_iterations_cheap_expensive"][0][0] == 21


def test_adaptive_solve_converges_to_the_direct_solution():
    dim = 3
    c0 = coeff.NSINKER_CENTERS[4]
    box = adaptive.AdaptiveBox.refined(dim, 2, lambda x, l: np.linalg.norm(x - c0) < 0.3, 1)
    assert box.n_cells == 106 and len(box.hanging_u) > 0
    uni = adaptive.AdaptiveBox.refined(dim, 2, lambda x, l: False, 0)
    its = {}
    for name, b in (("adaptive", box), ("uniform", uni)):
        eta = np.full(b.n_cells, 0.01)
        ops = b.assemble(eta, 1.0, "zero")
        rhs = np.concatenate([ops["Cu"].T @ np.sin(np.arange(b.n_u)), np.zeros(b.n_p)])
        out = adaptive.block_gmg_solve(b, eta, rhs, "zero", 1.0, 1e-8)
        assert out["converged"]
        its[name] = out["iterations"]
        if name == "adaptive":
            # direct solve of the same constrained system (one free pressure DoF pinned: the pressure has a constant null space)
            nu = b.n_u
            K = ops["K"].tolil()
            rb = rhs.copy(); rb[ops["constrained_u"]] = 0.0; rb[nu + ops["constrained_p"]] = 0.0
            free_p = min(set(range(b.n_p)) - set(ops["constrained_p"].tolist()))
            K[nu + free_p, :] = 0.0; K[nu + free_p, nu + free_p] = 1.0; rb[nu + free_p] = 0.0
            xd = spla.spsolve(K.tocsc(), rb)
            assert np.linalg.norm(out["x"][:nu] - xd[:nu]) < 1e-6 * np.linalg.norm(xd[:nu])
    # the refinement edge does not spoil the multigrid: same iteration count as the uniform mesh up to a few
    assert abs(its["adaptive"] - its["uniform"]) <= 5, its


@pytest.mark.parametrize("which", [1, 2, 3])
def test_muparser_temperature_example_golden_pins_the_adaptive_path(which):
    """tests/muparser_temperature_example of the reference: block GMG (local smoothing) on a 5 x 1-tree box, uniform and after two
    adaptive steps driven by a `minimum refinement function` alone -- the meshes are known without any error indicator, so this
    golden pins hanging nodes, refinement edges, interface operators and copy indices of the restatement against the reference:
    cells, levels and pressure DoFs exactly, velocity DoFs up to deal.II's identity-constrained duplicates on coarse-fine edges,
    iterations 6 / 9 / 9 within +-1 (7 / 9 / 10)."""
    from oracle import muparser_example as mx
    from util import golden
    g = golden("muparser_temperature_example")
    box, out = mx.solve(which)
    k = which - 1
    assert box.n_cells == g["active_cells"][k] and box.Lmax + 1 == g["levels"][k]
    assert box.n_p == g["dofs_u_p"][k][1]
    n_edges = len(box.hanging_p)                      # one hanging Q1 node per coarse-fine edge (2D)
    assert box.n_u + 2 * n_edges == g["dofs_u_p"][k][0], (box.n_u, n_edges)
    assert (n_edges == 0) == (which == 1)
    assert out["converged"] and abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][k][0]) <= 1, out["iterations"]


def test_initial_global_refinement_golden_anisotropic_cells_and_mixed_boundaries():
    """tests/initial_global_refinement of the reference: 1,855 cells of FIVE sizes (0.24 x 0.2 coarse cells, levels 0 .. 4 by depth),
    tangential velocity on one face and zero velocity on the other three, default solver (block GMG): the mesh is reproduced
    exactly (cells, levels, pressure DoFs; velocity DoFs up to deal.II's duplicates on the 75 coarse-fine edges), 23+0 iterations
    as 24."""
    from oracle import initial_global_refinement as igr
    g = golden("initial_global_refinement")
    box, out = igr.solve()
    assert box.n_cells == g["active_cells"][0] == 1855 and box.Lmax + 1 == g["levels"][0]
    assert sorted({c[0] for c in box.cells}) == [0, 1, 2, 3, 4]
    assert box.n_p == g["dofs_u_p"][0][1] and box.n_u + 2 * len(box.hanging_p) == g["dofs_u_p"][0][0]
    assert out["converged"] and abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][0][0]) <= 1, out["iterations"]
