"""Pin the oracle against the reference's own golden outputs (CPU only).

Goldens: /root/reference/tests/{sol_cx_2_gmg,sol_cx_4_gmg,sol_cx_4_gmg_no_avg,nsinker_gmg_gc,nsinker_gmg_no_avg}/screen-output
(deal.II 9.6; extracted once into tests/golden/reference_screen_output.json by tools/extract_goldens.py -- nothing is
read from /root/reference at run time).
The assembled-matrix backend is independent of any matrix-free code; the C matrix-free backend is then
checked against it operator by operator.
"""
import os

import numpy as np
import pytest

from oracle import assembled, coeff, cpu, mesh, postprocess, solver
from util import golden

HAVE_SOLCX_REF = os.path.exists(os.path.join(os.path.dirname(postprocess.__file__), "_ref", "libsolcx.so"))


def _nsinker(R, avg, backend, n_sinkers=4):
    hier = mesh.Hierarchy(3, R)
    act = hier.active
    eta_q, rho_q = coeff.nsinker(act.quadrature_points(), n_sinkers, 1e4)
    active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, avg)
    be = backend(hier, active, levels, 1.0, "zero")
    f = assembled.rhs_gravity(act, rho_q, np.array([0, 0, -1.0]))
    f[act.constrained_u("zero")] = 0
    rhs = np.concatenate([f, np.zeros(act.n_p)])
    out = solver.BlockGMGStokes(be).solve(rhs, tol_rel=1e-6, restart=100, max_its=1000)
    return act, emin, emax, out


@pytest.mark.parametrize("backend", [solver.AssembledBackend, cpu.backend_from_hierarchy], ids=["assembled", "matrix_free_c"])
def test_nsinker_gmg_goldens(backend):
    g = golden("nsinker_gmg_gc")                                           # mesh 1 of the test
    act, emin, emax, out = _nsinker(2, "harmonic", backend)
    assert [act.n_u, act.n_p] == g["dofs_u_p"][0] == [2187, 125]           # "2,187+125"
    assert [float(f"{emin:.6g}"), float(f"{emax:.6g}")] == g["viscosity_range"][0] == [0.01, 0.198921]   # "Viscosity range"
    assert out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == 21                         # "21+0"
    assert out["n_prec"] == g["A_preconditioner_applications"][0] == g["schur_preconditioner_applications"][0] == 22
    g = golden("nsinker_gmg_no_avg")
    act, emin, emax, out = _nsinker(2, "none", backend)
    assert [float(f"{emin:.6g}"), float(f"{emax:.6g}")] == g["viscosity_range"][0] == [0.01, 100.0]
    assert abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][0][0]) <= 1                      # "172+0" (+-1: App. D6)
    assert g["A_preconditioner_applications"][0] == 174 and out["n_prec"] == out["iterations"] + 2         # one extra per restart cycle


def _gold_solcx(R, avg):
    g = golden({(2, "harmonic"): "sol_cx_2_gmg", (4, "harmonic"): "sol_cx_4_gmg", (4, "none"): "sol_cx_4_gmg_no_avg"}[(R, avg)])
    return g["dofs_u_p"][0][0], g["dofs_u_p"][0][1], g["stokes_iterations_cheap_expensive"][0][0], ", ".join(g["errors_uL1_pL1_uL2_pL2"][0])


@pytest.mark.parametrize("R", [2, 4])
@pytest.mark.parametrize("avg", ["harmonic", "none"])
def test_solcx_goldens(R, avg):
    if avg == "none" and R == 2:
        pytest.skip("no golden for sol_cx_2 without averaging")
    hier = mesh.Hierarchy(2, R)
    act = hier.active
    eta_q, rho_q = coeff.solcx(act.quadrature_points())
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, avg)
    ps = coeff.pressure_scaling_box(hier, lambda x: coeff.solcx(x)[0])
    assert abs(ps - 1e5) < 1e-6
    be = solver.AssembledBackend(hier, active, levels, ps, "tangential")
    f = assembled.rhs_gravity(act, rho_q, np.array([0, -1.0]))
    f[be.cu[-1]] = 0
    rhs = np.concatenate([f, np.zeros(act.n_p)])
    s = solver.BlockGMGStokes(be)
    out = s.solve(rhs, tol_rel=1e-7, restart=100, max_its=1000)
    nu, npp, its, errs = _gold_solcx(R, avg)
    assert (nu, npp, its) == {2: (162, 25, 11), 4: (2178, 289, 15)}[R]
    assert (act.n_u, act.n_p) == (nu, npp)
    assert out["iterations"] == its and out["n_prec"] == its + 1
    assert s.solve(rhs, x0=out["x"], tol_rel=1e-7)["iterations"] == 0      # "0+0 iterations"
    if HAVE_SOLCX_REF:
        p, _ = postprocess.normalize_pressure_surface(act, out["x"][act.n_u:] * ps)
        e = postprocess.solcx_errors(act, out["x"][:act.n_u], p)
        assert ", ".join(f"{v:.6e}" for v in e) == errs


def test_idr_golden():
    """tests/nsinker_gmg_idr (R=2, 4 sinkers, harmonic average only viscosity, IDR(2)): "9+0 iterations", 25 applications of
    both preconditioners (8 complete IDR(2) steps x 3 + the early exit after the first inner step of the ninth)."""
    g = golden("nsinker_gmg_idr")
    hier = mesh.Hierarchy(3, 2)
    act = hier.active
    eta_q, rho_q = coeff.nsinker(act.quadrature_points(), 4, 1e4)
    active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, "harmonic")
    assert [float(f"{emin:.6g}"), float(f"{emax:.6g}")] == g["viscosity_range"][0]
    be = solver.AssembledBackend(hier, active, levels, 1.0, "zero")
    f = assembled.rhs_gravity(act, rho_q, np.array([0, 0, -1.0]))
    f[act.constrained_u("zero")] = 0
    rhs = np.concatenate([f, np.zeros(act.n_p)])
    out = solver.BlockGMGStokes(be).solve(rhs, tol_rel=1e-6, krylov="idr", idr_s=2, max_its=1000)
    assert out["converged"]
    assert out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == 9
    assert out["n_prec"] == g["A_preconditioner_applications"][0] == g["schur_preconditioner_applications"][0] == 25


def test_libstdcxx_normal_stream():
    """First values of std::mt19937 + std::normal_distribution<double> (libstdc++), the stream deal.II's SolverIDR draws its
    shadow vectors from; compared with a C++ program when g++ is available, else with recorded values."""
    import shutil, subprocess, tempfile
    got = solver.libstdcxx_normal_stream(7)
    recorded = None
    if shutil.which("g++"):
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "n.cc")
            open(src, "w").write('#include <random>\n#include <cstdio>\nint main(){std::mt19937 r;std::normal_distribution<double> n(0.0,1.0);'
                                 'for(int i=0;i<7;++i)std::printf("%.17g\\n",n(r));}\n')
            subprocess.check_call(["g++", "-O1", "-o", os.path.join(d, "n"), src])
            recorded = np.array([float(x) for x in subprocess.check_output([os.path.join(d, "n")]).split()])
        assert np.array_equal(got, recorded)
    assert got[0] == 0.13452965847232812


def test_no_cheap_gmg_golden():
    """tests/no_cheap_gmg (SolCx R=4, `harmonic average` of viscosity AND density, 0 cheap steps): the expensive phase
    (FGMRES + inner CG solves) takes "0+12 iterations" and the four SolCx errors match digit for digit."""
    g = golden("no_cheap_gmg")
    hier = mesh.Hierarchy(2, 4)
    act = hier.active
    eta_q, rho_q = coeff.solcx(act.quadrature_points())
    rho_q = coeff.average(rho_q, "harmonic")
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, "harmonic")
    ps = coeff.pressure_scaling_box(hier, lambda x: coeff.solcx(x)[0])
    be = solver.AssembledBackend(hier, active, levels, ps, "tangential")
    f = assembled.rhs_gravity(act, rho_q, np.array([0, -1.0]))
    f[be.cu[-1]] = 0
    rhs = np.concatenate([f, np.zeros(act.n_p)])
    out = solver.BlockGMGStokes(be).solve(rhs, tol_rel=1e-7, restart=100, max_its=0, n_expensive=1000)
    assert out["converged"]
    assert [out["iterations"], out["iterations_expensive"]] == g["stokes_iterations_cheap_expensive"][0] == [0, 12]
    if HAVE_SOLCX_REF:
        p, _ = postprocess.normalize_pressure_surface(act, out["x"][act.n_u:] * ps)
        e = postprocess.solcx_errors(act, out["x"][:act.n_u], p)
        assert [f"{v:.6e}" for v in e] == g["errors_uL1_pL1_uL2_pL2"][0]


@pytest.mark.parametrize("dim,R,bc", [(2, 3, "tangential"), (3, 2, "zero"), (3, 2, "tangential")])
def test_matrix_free_c_equals_assembled(dim, R, bc):
    rng = np.random.default_rng(0)
    hier = mesh.Hierarchy(dim, R)
    act = hier.active
    eta_q = np.exp(rng.uniform(-3, 3, size=(act.n_cells, 3 ** dim)))
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    for avg in ("none", "harmonic"):
        active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, avg)
        ab = solver.AssembledBackend(hier, active, levels, 3.7, bc)
        cb = cpu.backend_from_hierarchy(hier, active, levels, 3.7, bc)
        xu, xp = rng.standard_normal(act.n_u), rng.standard_normal(act.n_p)
        yu, yp = cb.stokes(xu, xp)
        zu, zp = ab.stokes(xu, xp)
        errs = [rel(yu, zu), rel(yp, zp), rel(cb.BT(xp), ab.BT(xp)), rel(cb.A_active(xu), ab.A_active(xu)), rel(cb.M_active(xp), ab.M_active(xp))]
        for l in range(hier.R + 1):
            L = hier.levels[l]
            x, q = rng.standard_normal(L.n_u), rng.standard_normal(L.n_p)
            errs += [rel(cb.apply_A(l, x), ab.apply_A(l, x)), rel(cb.apply_M(l, q), ab.apply_M(l, q)),
                     rel(cb.dinv_A(l), ab.dinv_A(l)), rel(cb.dinv_M(l), ab.dinv_M(l))]
            if l > 0:
                xc, qc = rng.standard_normal(hier.levels[l - 1].n_u), rng.standard_normal(hier.levels[l - 1].n_p)
                errs += [rel(cb.prolongate_u(l, xc), ab.prolongate_u(l, xc)), rel(cb.prolongate_p(l, qc), ab.prolongate_p(l, qc)),
                         rel(cb.restrict_u(l, x), ab.restrict_u(l, x)), rel(cb.restrict_p(l, q), ab.restrict_p(l, q))]
        assert max(errs) < 1e-13


def test_operator_properties():
    """Size-independent properties: symmetry of A and M, B^T adjoint to the pressure rows of the Stokes operator,
    restriction = transpose of prolongation, prolongation reproduces Q2/Q1 functions."""
    rng = np.random.default_rng(1)
    hier = mesh.Hierarchy(3, 2)
    act = hier.active
    eta_q = np.exp(rng.uniform(-3, 3, size=(act.n_cells, 27)))
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, "none")
    cb = cpu.backend_from_hierarchy(hier, active, levels, 2.0, "zero")
    x, y = rng.standard_normal(act.n_u), rng.standard_normal(act.n_u)
    assert abs(y @ cb.A_active(x) - x @ cb.A_active(y)) < 1e-10 * abs(y @ cb.A_active(x))
    p, q = rng.standard_normal(act.n_p), rng.standard_normal(act.n_p)
    assert abs(q @ cb.M_active(p) - p @ cb.M_active(q)) < 1e-12 * abs(q @ cb.M_active(p))
    xz = x.copy(); xz[act.constrained_u("zero")] = 0
    _, bp = cb.stokes(xz, np.zeros(act.n_p))
    assert abs(p @ bp - xz @ cb.BT(p)) < 1e-10 * abs(p @ bp)
    xc = rng.standard_normal(hier.levels[1].n_u)
    assert abs(x @ cb.prolongate_u(2, xc) - xc @ cb.restrict_u(2, x)) < 1e-12 * np.linalg.norm(x) * np.linalg.norm(xc)
    # a Q2 function on the coarse level is reproduced exactly on the fine level (no constraints)
    cb0 = cpu.backend_from_hierarchy(hier, active, levels, 2.0, "none")
    fc = lambda X: X[:, 0] ** 2 - 0.5 * X[:, 1] * X[:, 2] + X[:, 2] ** 2
    uc = fc(hier.levels[1].dof_coords_u())
    assert np.abs(cb0.prolongate_u(2, uc) - fc(hier.levels[2].dof_coords_u())).max() < 1e-14


@pytest.mark.parametrize("R,mesh_index", [(3, 0), (4, 1)])
def test_prescribed_dilation_gmg_golden(R, mesh_index):
    """tests/prescribed_dilation_gmg (2D box [0,2]^2, both meshes are uniform: refinement fraction 1): "35+0" / "36+0"
    iterations and the printed errors u_L2, p_L2 = 1.794305e-04, 8.789932e-05 / 2.245432e-05, 5.910004e-06.  Pins the
    prescribed-dilation branch (deviatoric active operators, right-hand side term -ps R q) AND that the level operators stay
    non-deviatoric (local_smoothing.cc:312); with deviatoric level operators the solve takes 26 / 27 iterations."""
    from oracle import prescribed_dilation as pd
    g = golden("prescribed_dilation_gmg")
    hier = mesh.Hierarchy(2, R, extents=[pd.EXTENT, pd.EXTENT])
    act = hier.active
    assert [act.n_u, act.n_p] == g["dofs_u_p"][mesh_index]
    eta_q, ps, rhs, u0 = pd.problem(act)
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, "harmonic")
    be = solver.AssembledBackend(hier, active, levels, ps, "zero", compressible=False)            # level operators: not deviatoric
    be.K, be.A_act, _ = assembled.stokes_matrix(act, active, ps, be.cu[-1], True)                 # active operators: deviatoric
    K_plain, _, _ = assembled.stokes_matrix(act, active, ps, None, True)
    rhs = rhs - K_plain @ np.concatenate([u0, np.zeros(act.n_p)])                                # correct_stokes_rhs
    rhs[be.cu[-1]] = 0.0
    out = solver.BlockGMGStokes(be).solve(rhs, tol_rel=1e-12, restart=100, max_its=200)
    assert out["converged"] and out["iterations"] == g["stokes_iterations_cheap_expensive"][2 * mesh_index][0]
    u = out["x"][:act.n_u] + u0
    p, _ = postprocess.normalize_pressure_volume(act, out["x"][act.n_u:] * ps)
    eu, ep = pd.errors(act, u, p)
    gu, gp = (float(v) for v in g["errors_uL2_pL2"][mesh_index])
    assert abs(eu / gu - 1) < 2e-6 and abs(ep / gp - 1) < 2e-6, (eu, ep, gu, gp)


def test_bicgstab_inner_solver():
    """The BiCGStab restatement (deal.II 9.6 SolverBicgstab, exact_residual = false) used for a non-symmetric A block in the
    expensive phase: solves the A block to the same solution as CG, and the outer FGMRES takes the same steps with either."""
    hier = mesh.Hierarchy(2, 4)
    act = hier.active
    eta_q, rho_q = coeff.solcx(act.quadrature_points())
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, "harmonic")
    ps = coeff.pressure_scaling_box(hier, lambda x: coeff.solcx(x)[0])
    be = solver.AssembledBackend(hier, active, levels, ps, "tangential")
    s = solver.BlockGMGStokes(be)
    b = np.sin(np.arange(act.n_u)); b[be.cu[-1]] = 0.0
    # (deal.II's breakdown threshold is the absolute value 1e-10 on r . rbar ~ |r|^2: tolerances below 1e-5 only restart)
    x1, it1, ok1 = solver.cg(be.A_active, s.mg_A.vmult, b, 1e-4 * np.linalg.norm(b), 1000)
    x2, it2, ok2 = solver.bicgstab(be.A_active, s.mg_A.vmult, b, 1e-4 * np.linalg.norm(b), 1000)
    assert ok1 and ok2 and np.linalg.norm(x1 - x2) < 1e-3 * np.linalg.norm(x1) and it2 <= it1
    # expensive phase with either inner solver on a problem of O(1) scaling (with SolCx's pressure scaling of 1e5 the residuals
    # of the A block fall below deal.II's absolute breakdown threshold and BiCGStab only restarts)
    rhs = np.concatenate([b, np.zeros(act.n_p)])
    outs = []
    for sym in (True, False):
        s.A_block_is_symmetric = sym
        outs.append(s.solve(rhs, tol_rel=1e-6, max_its=0, n_expensive=100))
    assert all(o["converged"] for o in outs) and abs(outs[0]["iterations_expensive"] - outs[1]["iterations_expensive"]) <= 1
    nu = act.n_u
    assert np.linalg.norm(outs[0]["x"][:nu] - outs[1]["x"][:nu]) < 1e-4 * np.linalg.norm(outs[0]["x"][:nu])
