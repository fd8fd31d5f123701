"""Pin the C++ host-core implementation of the solve (oracle/csrc/smf_cpu_solver.cc; test infrastructure, the CPU arm of
bench.py and the checker of the cfg 1 / cfg 2 GPU parity tests) against (a) the numpy/assembled oracle, piece by piece, and
(b) the reference's own goldens (tests/golden/reference_screen_output.json).  CPU only."""
import os

import numpy as np
import pytest

from oracle import assembled, coeff, cpuref, mesh, postprocess, solver
from util import golden, rel, seeded

HAVE_SOLCX_REF = os.path.exists(os.path.join(os.path.dirname(postprocess.__file__), "_ref", "libsolcx.so"))


def _problem(dim, R, bc, avg, n_sinkers=4, eta_per_q=False):
    hier = mesh.Hierarchy(dim, R)
    act = hier.active
    if dim == 3:
        eta_q, rho_q = coeff.nsinker(act.quadrature_points(), n_sinkers, 1e4)
        ps, g = 1.0, np.array([0, 0, -1.0])
    else:
        eta_q, rho_q = coeff.solcx(act.quadrature_points())
        ps, g = coeff.pressure_scaling_box(hier, lambda x: coeff.solcx(x)[0]), np.array([0, -1.0])
    active, levels, emin, emax = coeff.viscosity_tables(hier, eta_q, avg)
    f = assembled.rhs_gravity(act, rho_q, g)
    f[act.constrained_u(bc)] = 0
    rhs = np.concatenate([f, np.zeros(act.n_p)])
    return hier, active, levels, ps, rhs, (emin, emax)


@pytest.mark.parametrize("dim,R,bc,avg", [(3, 2, "zero", "harmonic"), (3, 2, "tangential", "none"), (2, 3, "tangential", "harmonic"),
                                           (2, 3, "zero", "none")])
def test_cpp_operators_equal_assembled(dim, R, bc, avg):
    """Every operator, diagonal and transfer of the C++ implementation against the assembled sparse matrices (1e-13)."""
    hier, active, levels, ps, rhs, _ = _problem(dim, R, bc, avg)
    ab = solver.AssembledBackend(hier, active, levels, ps, bc)
    cb = cpuref.from_hierarchy(hier, active, levels, ps, bc)
    for l, lev in enumerate(hier.levels):
        xu, xp = seeded(lev.n_u, 10 + l), seeded(lev.n_p, 20 + l)
        assert rel(cb.apply_A(l, xu), ab.apply_A(l, xu)) < 1e-13
        assert rel(cb.apply_M(l, xp), ab.apply_M(l, xp)) < 1e-13
        assert rel(cb.dinv_A(l), ab.dinv_A(l)) < 1e-13
        assert rel(cb.dinv_M(l), ab.dinv_M(l)) < 1e-13
        if l > 0:
            xc, xpc = seeded(hier.levels[l - 1].n_u, 30 + l), seeded(hier.levels[l - 1].n_p, 40 + l)
            assert rel(cb.prolongate_u(l, xc), ab.prolongate_u(l, xc)) < 1e-13
            assert rel(cb.prolongate_p(l, xpc), ab.prolongate_p(l, xpc)) < 1e-13
            assert rel(cb.restrict_u(l, xu), ab.restrict_u(l, xu)) < 1e-13
            assert rel(cb.restrict_p(l, xp), ab.restrict_p(l, xp)) < 1e-13
    act = hier.active
    xu, xp = seeded(act.n_u, 1), seeded(act.n_p, 2)
    xu[act.constrained_u(bc)] = 0
    (yu, yp), (zu, zp) = cb.stokes(xu, xp), ab.stokes(xu, xp)
    assert rel(yu, zu) < 1e-13 and rel(yp, zp) < 1e-13
    assert rel(cb.BT(xp), ab.BT(xp)) < 1e-13
    assert rel(cb.A_active(xu), ab.A_active(xu)) < 1e-13 and rel(cb.M_active(xp), ab.M_active(xp)) < 1e-13
    cb.close()


def test_cpp_smoothers_vcycles_preconditioner_equal_python():
    """Eigenvalue estimates (CG iteration counts and lambda_max per level), both V-cycles and the block preconditioner of the
    C++ implementation against oracle/solver.py over the assembled matrices."""
    hier, active, levels, ps, rhs, _ = _problem(3, 2, "zero", "harmonic")
    ab = solver.AssembledBackend(hier, active, levels, ps, "zero")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    s = solver.BlockGMGStokes(ab)
    cb.setup_smoothers()
    for l in range(hier.R + 1):
        for fam, mg in ((0, s.mg_A), (1, s.mg_S)):
            info = cb.smoother_info(l, fam)
            assert info["cg_iterations"] == mg.smoothers[l].cg_iterations
            assert abs(info["lambda_max"] / mg.smoothers[l].max_eig_est - 1) < 1e-10
            assert abs(info["theta"] / mg.smoothers[l].theta - 1) < 1e-10
    act = hier.active
    ru, rp = seeded(act.n_u, 5), seeded(act.n_p, 6)
    ru[act.constrained_u("zero")] = 0
    assert rel(cb.vcycle(0, ru), s.mg_A.vmult(ru)) < 1e-10
    assert rel(cb.vcycle(1, rp), s.mg_S.vmult(rp)) < 1e-10
    r = np.concatenate([ru, rp])
    assert rel(cb.precondition(r), s.prec(r)) < 1e-10
    cb.close()


def test_cpp_solve_reproduces_reference_goldens():
    """nsinker_gmg_gc: 21+0 iterations, 22 preconditioner applications; nsinker_gmg_no_avg: 172 (+-1), one extra application
    per restart cycle -- with the whole solver loop in C++; the solution equals the numpy-driven oracle's to 1e-10."""
    hier, active, levels, ps, rhs, rng_ = _problem(3, 2, "zero", "harmonic")
    g = golden("nsinker_gmg_gc")
    assert [float(f"{rng_[0]:.6g}"), float(f"{rng_[1]:.6g}")] == g["viscosity_range"][0]
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    out = cb.solve(rhs, tol_rel=1e-6, restart=100, max_its=1000)
    assert out["converged"] and out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == 21
    assert out["n_prec"] == g["A_preconditioner_applications"][0] == 22
    ref = solver.BlockGMGStokes(solver.AssembledBackend(hier, active, levels, ps, "zero")).solve(rhs, tol_rel=1e-6, restart=100, max_its=1000)
    nu = hier.active.n_u                        # (the pressure is only defined up to a constant on the no-slip box)
    assert ref["iterations"] == out["iterations"] and rel(out["x"][:nu], ref["x"][:nu]) < 1e-10
    assert abs(out["tolerance"] / ref["tolerance"] - 1) < 1e-12
    cb.close()
    hier, active, levels, ps, rhs, _ = _problem(3, 2, "zero", "none")
    g = golden("nsinker_gmg_no_avg")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    out = cb.solve(rhs, tol_rel=1e-6, restart=100, max_its=1000)
    assert abs(out["iterations"] - g["stokes_iterations_cheap_expensive"][0][0]) <= 1                  # 172
    assert out["n_prec"] == out["iterations"] + 2 and g["A_preconditioner_applications"][0] == 174
    cb.close()


@pytest.mark.parametrize("R,its", [(2, 11), (4, 15)])
def test_cpp_solcx_goldens(R, its):
    """sol_cx_2_gmg / sol_cx_4_gmg: iterations and the four error norms digit for digit, solved by the C++ loop (2D kernels,
    tangential boundaries, pressure scaling 1e5)."""
    hier, active, levels, ps, rhs, _ = _problem(2, R, "tangential", "harmonic")
    g = golden(f"sol_cx_{R}_gmg")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "tangential")
    out = cb.solve(rhs, tol_rel=1e-7, restart=100, max_its=1000)
    assert out["converged"] and out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == its
    assert out["n_prec"] == its + 1
    if HAVE_SOLCX_REF:
        act = hier.active
        p, _ = postprocess.normalize_pressure_surface(act, out["x"][act.n_u:] * ps)
        e = postprocess.solcx_errors(act, out["x"][:act.n_u], p)
        assert ", ".join(f"{v:.6e}" for v in e) == ", ".join(g["errors_uL1_pL1_uL2_pL2"][0])
    cb.close()


def test_cpp_backend_drives_python_idr_golden():
    """tests/nsinker_gmg_idr (9 iterations, 25 preconditioner applications) with the Python restatement of IDR(2) on top of
    the C++ operators (the combination tests/test_gpu_solve.py uses as checker at refinement 6)."""
    hier, active, levels, ps, rhs, _ = _problem(3, 2, "zero", "harmonic")
    g = golden("nsinker_gmg_idr")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    out = solver.BlockGMGStokes(cb).solve(rhs, tol_rel=1e-6, krylov="idr", idr_s=2, max_its=1000)
    assert out["converged"] and out["iterations"] == g["stokes_iterations_cheap_expensive"][0][0] == 9
    assert out["n_prec"] == g["A_preconditioner_applications"][0] == 25
    cb.close()


def test_cpp_simd_width_and_threads_do_not_change_the_result():
    """4-cell (AVX2) and 8-cell (AVX-512, if the host has it) batches, 1 thread and all threads: same iterations, solution
    equal to round-off (the colouring fixes the order of the additions into a DoF, so runs are reproducible)."""
    hier, active, levels, ps, rhs, _ = _problem(3, 2, "zero", "harmonic")
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    n_all = cpuref.set_num_threads()
    a = cb.solve(rhs, tol_rel=1e-6)
    b = cb.solve(rhs, tol_rel=1e-6)
    assert a["iterations"] == b["iterations"] and np.array_equal(a["x"], b["x"])
    cpuref.set_num_threads(1)
    c = cb.solve(rhs, tol_rel=1e-6)
    cpuref.set_num_threads(n_all)
    nu = hier.active.n_u
    assert c["iterations"] == a["iterations"] and rel(c["x"][:nu], a["x"][:nu]) < 1e-10
    cb.close()
