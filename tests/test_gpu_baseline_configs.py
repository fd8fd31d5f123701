"""GPU parity at the sizes BASELINE.json names, against the C++ host-core implementation of the reference algorithm
(oracle/csrc/smf_cpu_solver.cc, pinned by tests/test_cpuref.py) run on the GPU box's own cores in the same test:

  cfg 2  benchmarks/nsinker 3D, 8 sinkers, DR 1e4, global refinement 5   iterations equal, velocity rel. L2 <= 1e-8
  cfg 1  benchmarks/solcx 2D, global refinement 6                          iterations equal, velocity / pressure <= 1e-8
  R = 6  nsinker (performance_nsinker_gmg_idr mesh, GMRES)                 iterations equal, velocity <= 1e-8
  R = 7  the size bench.py times: the committed fixture of ONE complete CPU solve (tools/make_full_size_fixture.py; 70
         iterations, norms and 2 x 4096 seeded samples of the solution) -- the oracle is not re-run at that size here.
Acceptance (BASELINE.json north_star): outer iterations within +-1, velocity / pressure relative L2 <= 1e-8."""
import os

import numpy as np
import pytest

from oracle import coeff, cpuref, postprocess
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import rel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gpu(dim, R, bc, problem, avg, tol, **kw):
    hier = synthetic.BoxHierarchy(dim, R)
    eta_q, rhs, ps = problem(hier, averaging=avg, **kw)
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, avg, ps)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": tol, "Number of cheap Stokes solver steps": 200})
    return hier, h, eta_q, rhs, ps


def _cpu(hier, eta_q, ps, bc, rhs, tol):
    active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, "harmonic")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, bc)
    cpuref.set_num_threads()
    out = cb.solve(rhs, tol_rel=tol, restart=100, max_its=200)
    cb.close()
    return out, (emin, emax)


@pytest.mark.parametrize("R,n_sinkers,avg", [(5, 8, "harmonic average"), (6, 4, "harmonic average only viscosity")], ids=["cfg2_R5", "R6"])
def test_nsinker_matches_cpu_reference(R, n_sinkers, avg):
    hier, h, eta_q, rhs, ps = _gpu(3, R, "zero", benchmarks.nsinker_problem, avg, 1e-6, n_sinkers=n_sinkers)
    sol, out = h.solve(rhs)
    ref, (emin, emax) = _cpu(hier, eta_q, ps, "zero", rhs, 1e-6)
    assert out["converged"] and ref["converged"]
    assert abs(h.minimum_viscosity / emin - 1) < 1e-12 and abs(h.maximum_viscosity / emax - 1) < 1e-12
    assert out["iterations"] == ref["iterations"], (out["iterations"], ref["iterations"])
    assert out["n_A_preconditioner"] == ref["n_prec"]
    assert abs(out["tolerance"] / ref["tolerance"] - 1) < 1e-10
    nu = hier.active.n_u
    assert rel(sol[:nu], ref["x"][:nu]) < 1e-8
    pg, po = sol[nu:] - sol[nu:].mean(), ref["x"][nu:] * ps - (ref["x"][nu:] * ps).mean()      # no-slip box: p up to a constant
    assert rel(pg, po) < 1e-8
    h.close()


def test_solcx_r6_matches_cpu_reference():
    """BASELINE cfg 1: SolCx at global refinement 6 (33,282 + 4,225 DoFs), tangential boundaries, pressure scaling 1e5."""
    avg = "harmonic average only viscosity"
    hier, h, eta_q, rhs, ps = _gpu(2, 6, "tangential", benchmarks.solcx_problem, avg, 1e-7)
    assert (hier.active.n_u, hier.active.n_p) == (33282, 4225)
    sol, out = h.solve(rhs)
    ref, _ = _cpu(hier, eta_q, ps, "tangential", rhs, 1e-7)
    assert out["converged"] and out["iterations"] == ref["iterations"]
    nu = hier.active.n_u
    assert rel(sol[:nu], ref["x"][:nu]) < 1e-8
    pg, po = sol[nu:] - sol[nu:].mean(), ref["x"][nu:] * ps - (ref["x"][nu:] * ps).mean()
    assert rel(pg, po) < 1e-8
    # the four analytic errors agree with the CPU implementation's to the digits the reference prints
    if os.path.exists(os.path.join(os.path.dirname(postprocess.__file__), "_ref", "libsolcx.so")):
        from oracle import mesh as omesh
        act = omesh.Hierarchy(2, 6).active
        eg = postprocess.solcx_errors(act, sol[:nu], postprocess.normalize_pressure_surface(act, sol[nu:])[0])
        ec = postprocess.solcx_errors(act, ref["x"][:nu], postprocess.normalize_pressure_surface(act, ref["x"][nu:] * ps)[0])
        assert [f"{v:.6e}" for v in eg] == [f"{v:.6e}" for v in ec]
    h.close()


def test_refinement7_matches_cpu_fixture():
    """The configuration bench.py times (53 M DoFs, 8 levels) against the committed complete CPU solve."""
    fx = np.load(os.path.join(GOLD, "nsinker_R7_cpu_reference.npz"))
    hier, h, eta_q, rhs, ps = _gpu(3, 7, "zero", benchmarks.nsinker_problem, "harmonic average", 1e-6, n_sinkers=8)
    nu, npr = hier.active.n_u, hier.active.n_p
    assert (nu, npr) == (int(fx["n_u"]), int(fx["n_p"]))
    assert abs(np.linalg.norm(rhs) / float(fx["rhs_norm"]) - 1) < 1e-12
    assert abs(h.minimum_viscosity / fx["viscosity_range"][0] - 1) < 1e-12 and abs(h.maximum_viscosity / fx["viscosity_range"][1] - 1) < 1e-12
    sol, out = h.solve(rhs)
    assert out["converged"]
    assert abs(out["iterations"] - int(fx["iterations"])) <= 1, (out["iterations"], int(fx["iterations"]))
    assert abs(out["tolerance"] / float(fx["tolerance"]) - 1) < 1e-9
    u, p = sol[:nu], sol[nu:] - sol[nu:].mean()
    if out["iterations"] == int(fx["iterations"]):
        assert abs(np.linalg.norm(u) / float(fx["u_norm"]) - 1) < 1e-8
        assert rel(u[fx["u_index"]], fx["u_sample"]) < 1e-8
        assert rel(p[fx["p_index"]], fx["p_sample"]) < 1e-7
    else:          # one iteration more or less: both are converged to 1e-6, the iterates differ at that level
        assert rel(u[fx["u_index"]], fx["u_sample"]) < 1e-4
    h.close()
