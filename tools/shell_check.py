"""GPU diagnostics for the curved-cell path (spherical shell): prints the operator errors against the oracle at refinement 1,
the solve comparison, and kernel timings at a larger refinement.  python tools/shell_check.py [R_timing]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from aspect_b200 import StokesMatrixFreeHandler, SMF_ACTIVE, benchmarks          # noqa: E402
from aspect_b200 import shell as pshell                                           # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def parity():
    from oracle import assembled, coeff, solver
    from oracle import shell as oshell
    ohier, hier = oshell.ShellHierarchy(1), pshell.ShellHierarchy(1)
    act = hier.active
    rng = np.random.default_rng(5)
    eta_q = np.exp(rng.uniform(-4, 4, size=(act.n_cells, 27)))
    xu, xp = rng.uniform(-1, 1, act.n_u), rng.uniform(-1, 1, act.n_p)
    for averaging in ("none", "harmonic average"):
        op = {"none": "none", "harmonic average": "harmonic"}[averaging]
        active, levels, emin, emax = coeff.viscosity_tables(ohier, eta_q, op)
        ob = solver.AssembledBackend(ohier, active, levels, 3.7, "zero")
        for fast in (True, False):
            h = StokesMatrixFreeHandler(3)
            h.curved_sum_factorization = fast
            h.setup_dofs(hier, "zero")
            h.assemble(eta_q, averaging, 3.7)
            h.build_preconditioner()
            src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
            src[0].from_host(xu); src[1].from_host(xp)
            h.stokes_matrix.vmult(dst, src)
            yu, yp = ob.stokes(xu, xp)
            e = {"stokes_u": rel(dst[0].to_host(), yu), "stokes_p": rel(dst[1].to_host(), yp)}
            h.BT_block.vmult(dst, src); e["BT"] = rel(dst[0].to_host(), ob.BT(xp))
            h.A_block_matrix.vmult(dst[0], src[0]); e["A"] = rel(dst[0].to_host(), ob.A_active(xu))
            h.Schur_complement_block_matrix.vmult(dst[1], src[1]); e["M"] = rel(dst[1].to_host(), ob.M_active(xp))
            e["dinvA"] = rel(h.get_matrix_diagonal_inverse(SMF_ACTIVE, 0), 1.0 / ob.A_act.diagonal())
            e["dinvM"] = rel(h.get_matrix_diagonal_inverse(SMF_ACTIVE, 1), 1.0 / ob.M_act.diagonal())
            e["table"] = rel(h.get_cell_data(SMF_ACTIVE), active)
            print("parity", averaging, "fast" if fast else "general", {k: "%.1e" % v for k, v in e.items()}, flush=True)
            h.close()
    # solve
    eta_q, force_q, ps = benchmarks.nsinker_shell_problem(hier)
    L = ohier.active
    eo, ro = coeff.nsinker_shell(L.xq)
    active, levels, emin, emax = coeff.viscosity_tables(ohier, eo, "harmonic")
    g = -L.xq / np.linalg.norm(L.xq, axis=-1, keepdims=True)
    f = assembled.rhs_force(L, coeff.average(ro, "harmonic")[..., None] * g)
    f[L.constrained_u("zero")] = 0.0
    ob = solver.AssembledBackend(ohier, active, levels, ps, "zero")
    t0 = time.time()
    want = solver.BlockGMGStokes(ob).solve(np.concatenate([f, np.zeros(L.n_p)]), tol_rel=1e-6, restart=200, max_its=2000)
    print("oracle solve: its", want["iterations"], "n_prec", want["n_prec"], "%.1fs" % (time.time() - t0), "range", emin, emax, "ps", ps, flush=True)
    for fast in (True, False):
        h = StokesMatrixFreeHandler(3)
        h.curved_sum_factorization = fast
        h.setup_dofs(hier, "zero")
        h.assemble(eta_q, "harmonic average", ps)
        h.build_preconditioner()
        h.parameters.update({"Linear solver tolerance": 1e-6, "Number of cheap Stokes solver steps": 2000,
                             "Maximum number of expensive Stokes solver steps": 0, "GMRES solver restart length": 200})
        r = h.initialize_dof_vector(0)
        h.assemble_stokes_rhs(force_q, r)
        fr = r.to_host()
        sol, out = h.solve(np.concatenate([fr, np.zeros(act.n_p)]))
        nu = act.n_u
        xs = want["x"].copy(); xs[nu:] *= ps
        print("gpu solve", "fast" if fast else "general", "its", out["iterations"], "nA", out["n_A_preconditioner"], "rhs err %.1e" % rel(fr, f),
              "u err %.2e" % rel(sol[:nu], xs[:nu]), "p err %.2e" % rel(sol[nu:] - sol[nu:].mean(), xs[nu:] - xs[nu:].mean()),
              "range", h.minimum_viscosity, h.maximum_viscosity, "solve_s %.3f" % out["solve_seconds"], flush=True)
        h.close()


def timing(R):
    t0 = time.time()
    hier = pshell.ShellHierarchy(R)
    eta_q, force_q, ps = benchmarks.nsinker_shell_problem(hier)
    act = hier.active
    print("shell R=%d: %d cells, %d + %d DoFs, host set-up %.1fs" % (R, act.n_cells, act.n_u, act.n_p, time.time() - t0), flush=True)
    for averaging in ("harmonic average", "none"):
        for fast in (True, False):
            h = StokesMatrixFreeHandler(3)
            h.curved_sum_factorization = fast
            h.setup_dofs(hier, "zero")
            h.assemble(eta_q, averaging, ps)
            h.build_preconditioner()
            src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
            rng = np.random.default_rng(0)
            src[0].from_host(rng.uniform(-1, 1, act.n_u)); src[1].from_host(rng.uniform(-1, 1, act.n_p))
            ops = {"stokes": lambda: h.stokes_matrix.vmult(dst, src), "A": lambda: h.A_block_matrix.vmult(dst[0], src[0]),
                   "BT": lambda: h.BT_block.vmult(dst, src), "mass": lambda: h.Schur_complement_block_matrix.vmult(dst[1], src[1])}
            res = {}
            for name, fn in ops.items():
                for _ in range(3):
                    fn()
                h.synchronize()
                n = 20 if fast else 5
                t = time.perf_counter()
                for _ in range(n):
                    fn()
                h.synchronize()
                res[name] = (time.perf_counter() - t) / n
            n_eta = 27 if averaging == "none" else 1
            alg = {"stokes": 16 * (act.n_u + act.n_p) + 8 * act.n_cells * (n_eta + 270), "A": 16 * act.n_u + 8 * act.n_cells * (n_eta + 270)}
            print("timing", averaging, "fast" if fast else "general",
                  {k: "%.3f ms" % (1e3 * v) for k, v in res.items()},
                  "stokes %.1f GDoF/s, %.0f GB/s algorithmic; A %.1f GDoF/s, %.0f GB/s" % (
                      (act.n_u + act.n_p) / res["stokes"] / 1e9, alg["stokes"] / res["stokes"] / 1e9, act.n_u / res["A"] / 1e9, alg["A"] / res["A"] / 1e9), flush=True)
            if fast and averaging == "harmonic average":
                h.parameters.update({"Linear solver tolerance": 1e-6, "Number of cheap Stokes solver steps": 2000,
                                     "Maximum number of expensive Stokes solver steps": 0, "GMRES solver restart length": 200})
                r = h.initialize_dof_vector(0)
                h.assemble_stokes_rhs(force_q, r)
                sol, out = h.solve(np.concatenate([r.to_host(), np.zeros(act.n_p)]))
                print("solve R=%d:" % R, "its", out["iterations"], "converged", out["converged"], "setup_s %.3f solve_s %.3f" % (out["setup_seconds"], out["solve_seconds"]),
                      "viscosity range", h.minimum_viscosity, h.maximum_viscosity, flush=True)
            h.close()


def profile(R):
    """A few launches of the curved-cell kernels for an ncu capture."""
    hier = pshell.ShellHierarchy(R)
    eta_q, force_q, ps = benchmarks.nsinker_shell_problem(hier)
    act = hier.active
    h = StokesMatrixFreeHandler(3)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    src = h.initialize_dof_vector(); dst = h.initialize_dof_vector()
    rng = np.random.default_rng(0)
    src[0].from_host(rng.uniform(-1, 1, act.n_u)); src[1].from_host(rng.uniform(-1, 1, act.n_p))
    for _ in range(3):
        h.A_block_matrix.vmult(dst[0], src[0])
        h.stokes_matrix.vmult(dst, src)
        h.BT_block.vmult(dst, src)
    h.synchronize()
    h.close()


if __name__ == "__main__":
    if "--profile" in sys.argv:
        profile(int(sys.argv[-1]))
        sys.exit(0)
    if "--no-parity" not in sys.argv:
        parity()
    Rs = [int(a) for a in sys.argv[1:] if a.isdigit()]
    for R in Rs:
        timing(R)
