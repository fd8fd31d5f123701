"""The product's adaptive-mesh driver (aspect_b200/adaptive.py, vectorised over lattice keys) against the oracle's independent
dictionary-based restatement (oracle/adaptive.py): identical index maps, hanging-node lines, refinement-edge indices, child
tables, copy indices and level viscosities.  CPU only."""
import numpy as np
import pytest

from oracle import adaptive as oad
from aspect_b200 import adaptive as pad

CASES = [(2, 1, 2, lambda x, l: np.linalg.norm(x) < 0.6), (2, 2, 2, lambda x, l: abs(x[0] - 0.5) < 0.2 and x[1] > 0.4),
         (3, 1, 2, lambda x, l: np.linalg.norm(x) < 0.6), (3, 2, 1, lambda x, l: np.linalg.norm(x - 0.55) < 0.3),
         (3, 1, 0, lambda x, l: False)]


@pytest.mark.parametrize("dim,base,n_adaptive,marker", CASES)
def test_driver_equals_oracle(dim, base, n_adaptive, marker):
    box = oad.AdaptiveBox.refined(dim, base, marker, n_adaptive)
    ph = pad.AdaptiveBoxHierarchy.refined(dim, base, marker, n_adaptive)
    _compare(box, ph, dim)


@pytest.mark.parametrize("which", [1, 2, 3])
def test_driver_equals_oracle_on_several_coarse_cells(which):
    """The three meshes of the reference's tests/muparser_temperature_example (5 x 1 coarse cells of 1,000 km, refined by depth)."""
    from oracle import muparser_example as mx
    box = mx.mesh(which)
    ph = pad.AdaptiveBoxHierarchy(2, sorted(box.leaves), extents=(5e6, 1e6), repetitions=(5, 1))
    _compare(box, ph, 2)


def _compare(box, ph, dim):
    oh = oad.AdaptiveHierarchy(box, "zero")
    act = ph.active
    # active mesh
    assert act.n_cells == box.n_cells and [(int(l), tuple(int(v) for v in c)) for l, c in zip(act.cell_level, act.cells)] == box.cells
    assert np.array_equal(act.cell_dofs_u, box.cell_dofs_u) and np.array_equal(act.cell_dofs_p, box.cell_dofs_p)
    assert (act.n_u, act.n_p) == (box.n_u, box.n_p)
    assert np.array_equal(act.constrained_u("zero"), box.dirichlet_u("zero"))
    assert np.array_equal(act.constrained_u("tangential"), box.dirichlet_u("tangential"))
    for mine, theirs in ((act.hanging_u, box.hanging_u), (act.hanging_p, box.hanging_p)):
        assert sorted(mine) == sorted(theirs)
        for k in mine:
            a = dict(zip(mine[k][0].tolist(), mine[k][1].tolist()))
            b = dict(theirs[k])
            assert a.keys() == b.keys() and max(abs(a[m] - b[m]) for m in a) < 1e-15
    # constraint matrices from the CSR lines
    Cu, Cp, cons_u, cons_p = box.constraint_matrices("zero")
    dof, start, master, weight = act.constraint_lines("u", "zero")
    hang = set(dof.tolist())
    assert hang | set(act.constrained_u("zero").tolist()) == set(cons_u.tolist())
    x = np.random.default_rng(0).uniform(-1, 1, act.n_u)
    y = x.copy(); y[act.constrained_u("zero").astype(np.int64)] = 0.0
    for i, k in enumerate(dof):
        y[k] = weight[start[i]:start[i + 1]] @ x[master[start[i]:start[i + 1]].astype(np.int64)]
    assert np.allclose(y, Cu @ x, atol=1e-15)
    dofp, startp, masterp, weightp = act.constraint_lines("p")
    assert np.array_equal(dofp, cons_p)
    # levels
    assert len(ph.levels) == oh.n_levels
    for l, (L, O) in enumerate(zip(ph.levels, oh.levels)):
        assert np.array_equal(L.cell_dofs_u, O["cell_dofs_u"]) and np.array_equal(L.cell_dofs_p, O["cell_dofs_p"])
        assert np.array_equal(L.constrained_u("zero"), O["boundary_u"])
        assert np.array_equal(L.edge_u, O["edge_u"]) and np.array_equal(L.edge_p, O["edge_p"])
        au, lu, ap, lp = ph.copy_indices(l)
        assert np.array_equal(au, oh.copy_u[l][0]) and np.array_equal(lu, oh.copy_u[l][1])
        assert np.array_equal(ap, oh.copy_p[l][0]) and np.array_equal(lp, oh.copy_p[l][1])
        if l > 0:
            ch = ph.children(l)
            for pk, par in enumerate(oh.cells[l - 1]):
                kids = oad._children(par, dim)
                want = [oh.index[l].get(k, -1) for k in kids]
                assert ch[pk].tolist() == want
    eta = np.random.default_rng(1).uniform(0.5, 2.0, box.n_cells)
    for a, b in zip(ph.level_viscosity(eta), oh.level_viscosity(eta)):
        assert np.allclose(a, b, rtol=1e-15)
    # every active DoF is copied to exactly one level, except hanging ones (they sit on a refinement edge and do not exist
    # one level down)
    seen = np.concatenate([ph.copy_indices(l)[0] for l in range(len(ph.levels))])
    assert len(seen) == len(set(seen.tolist()))
    hanging_dofs = {dim * k + c for k in act.hanging_u for c in range(dim)}
    assert set(range(act.n_u)) - set(seen.tolist()) <= hanging_dofs
    # geometry tables of the active mesh
    assert act.jinv.shape == (act.n_cells, 3 ** dim, dim, dim) and np.allclose(act.jxw.sum(), np.prod(ph.extents))
    assert np.allclose(act.quadrature_points(), box.quadrature_points())
