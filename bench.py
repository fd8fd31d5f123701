#!/usr/bin/env python
"""Benchmark of the B200-native block-GMG Stokes path (BASELINE.json metric:
"Stokes vmult GDoF/s + nsinker block-GMG Stokes solve wall-time, 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference algorithm

One JSON line on stdout (rank 0).  A "step" is ONE complete block-GMG Stokes solve (StokesMatrixFreeHandler::
solve(), /root/reference/source/simulator/solver/stokes_matrix_free_local_smoothing.cc:758-1313: Chebyshev set-up
+ eigenvalue estimation + tolerance + right-preconditioned GMRES(100) to 1e-6) of the synthetic nsinker problem
(benchmarks/nsinker/nsinker_gmg.prm: 8 sinkers, viscosity ratio 1e4, harmonic averaging, no-slip box).

  value        solve throughput = (N_u + N_p) / wall time of one solve, vectors resident in HBM        [MDoF/s]
  e2e          the same through the reference-facing host call (StokesMatrixFreeHandler.solve -> smf_solve_host):
               right-hand side and initial guess copied from pinned host memory, solution copied back, every step
  vmult        Stokes / A-block operator throughput in GDoF/s (CUDA events, 20 launches, vectors >> L2)
  roofline     finest-level A-block kernel (the dominant kernel of the solve) against the measured HBM peak
  cpu_baseline the oracle (CPU restatement, OpenMP) on the host cores, complete solve at a smaller refinement

Everything under oracle/ is used here only by the cpu_baseline leg and the --impl reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SINKERS, VISCOSITY_RATIO, AVERAGING = 8, 1e4, "harmonic average"
TOL, RESTART, CHEAP_STEPS = 1e-6, 100, 200          # nsinker_gmg.prm
METRIC = "nsinker_block_gmg_stokes_solve_throughput"
UNIT = "MDoF/s"


REPETITIONS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}      # coarse cells of the unit box, one tree per GPU


def workload_name(R, world=1):
    rep = REPETITIONS[world]
    mesh = f"global refinement {R}" if world == 1 else (
        f"{rep[0]}x{rep[1]}x{rep[2]} coarse cells x global refinement {R}" + (f" (= refinement {R + 1} of the unit cube)" if world == 8 else ""))
    return (f"nsinker 3D unit box, {N_SINKERS} sinkers, viscosity ratio 1e4, harmonic average, Q2-Q1, block GMG "
            f"(GMRES({RESTART}), tol {TOL:g}), {mesh}")


# ------------------------------------------------------------------------------------------------------------
# The CPU arm: oracle/csrc/smf_cpu_solver.cc (C++17, OpenMP over all host cores, SIMD cell batches, the whole solver loop in
# C++; kind "port" -- deal.II / ASPECT cannot be built in this image).  Complete solves up to refinement 5; above that a
# complete solve takes minutes, so a step is a BOUNDED SAMPLE of the same solve on the same mesh: smoother set-up + eigenvalue
# estimation + tolerance + the first SAMPLE_ITS GMRES iterations, and the wall time of the complete solve follows from the
# sample's own timers (every part is linear in what it repeats):
#     t_solve = t_sample_outside_the_iteration_loop + its * (t_iteration_without_orthogonalisation + t_ortho((its + 1) / 2))
# (the orthogonalisation of iteration j streams j basis vectors; its mean cost over a cycle is the cost at the mean basis
# length, measured with the solve's own routine on a basis of that length) where `its` is the iteration count of the CPU implementation's own complete solve of that configuration (refinement 7:
# tests/golden/nsinker_R7_cpu_reference.npz, made once by tools/make_full_size_fixture.py).  The model is checked in every run
# against complete solves at refinement 5 (`model_check`).
SAMPLE_ITS = 2


def _cpu_problem(R, hier=None, eta_q=None, rhs=None, ps=None, n_sinkers=N_SINKERS):
    from oracle import coeff, cpuref
    from aspect_b200 import benchmarks, synthetic
    if hier is None:
        hier = synthetic.BoxHierarchy(3, R)
        eta_q, rhs, ps = benchmarks.nsinker_problem(hier, n_sinkers, VISCOSITY_RATIO, AVERAGING)
    active, levels, _, _ = coeff.viscosity_tables(hier, eta_q, "harmonic")
    threads = cpuref.set_num_threads()                # all cores the process may use (torchrun exports OMP_NUM_THREADS=1)
    cb = cpuref.from_hierarchy(hier, active, levels, ps, "zero")
    return cb, np.ascontiguousarray(rhs), hier.active.n_u + hier.active.n_p, threads


def _model_time(out, its_full, t_ortho_mean):
    """Wall time of the complete solve from a sample of m = out['iterations'] iterations: everything outside the iteration loop
    (set-up, tolerance, initial residual, the final update with its preconditioner application) happens once in both; inside
    the loop everything but the orthogonalisation costs the same every iteration, and the orthogonalisation of iteration j
    streams j basis vectors, so its mean over a cycle is the cost at the mean basis length (measured: smfc_time_ortho)."""
    m = out["iterations"]
    sp = out["split"]
    t_outside = out["seconds"] - sp["iteration_loop"]
    t_iter = (sp["iteration_loop"] - sp["gmres_ortho"]) / m
    return t_outside + its_full * (t_iter + t_ortho_mean)


def cpu_reference_steps(R, steps, warmup, hier=None, eta_q=None, rhs=None, ps=None, its_full=None):
    """`warmup` + `steps` steps of the CPU arm at refinement R; returns the dict both arms print."""
    from oracle import cpuref
    cb, rhs, n, threads = _cpu_problem(R, hier, eta_q, rhs, ps)
    width = int(cpuref.lib().smfc_width())
    times, its = [], None
    if R <= 5:
        for i in range(warmup + steps):
            out = cb.solve(rhs, tol_rel=TOL, restart=RESTART, max_its=CHEAP_STEPS)
            its = out["iterations"]
            if i >= warmup:
                times.append(out["seconds"])
        t = float(np.mean(times))
        sample = (f"{len(times)} complete solve(s) of this configuration (refinement {R}, {n} DoFs, {its} outer iterations, {t:.2f} s each)")
        model_check = None
    else:
        if its_full is None:
            fx = os.path.join(ROOT, "tests", "golden", f"nsinker_R{R}_cpu_reference.npz")
            if os.path.exists(fx):
                its_full = int(np.load(fx)["iterations"])
            else:                                         # no fixture for this size: one complete solve, outside the timed steps
                its_full = cb.solve(rhs, tol_rel=TOL, restart=RESTART, max_its=CHEAP_STEPS)["iterations"]
        t_ortho = cb.time_ortho((its_full + 1) // 2)
        for i in range(warmup + steps):
            out = cb.solve(rhs, tol_rel=TOL, restart=RESTART, max_its=SAMPLE_ITS)
            if i >= warmup:
                times.append(_model_time(out, its_full, t_ortho))
        its = its_full
        t = float(np.mean(times))
        # the same model against complete solves where those are affordable (cfg 2)
        cb5, rhs5, n5, _ = _cpu_problem(5)
        full = [cb5.solve(rhs5, tol_rel=TOL, restart=RESTART, max_its=CHEAP_STEPS) for _ in range(2)][-1]
        part = [cb5.solve(rhs5, tol_rel=TOL, restart=RESTART, max_its=SAMPLE_ITS) for _ in range(2)][-1]
        model_check = {"refinement": 5, "complete_solve_s": full["seconds"],
                       "modelled_from_sample_s": _model_time(part, full["iterations"], cb5.time_ortho((full["iterations"] + 1) // 2)),
                       "iterations": full["iterations"]}
        cb5.close()
        sample = (f"{len(times)} bounded sample(s) of this configuration (refinement {R}, {n} DoFs): smoother set-up + eigenvalue estimation + "
                  f"tolerance + the first {SAMPLE_ITS} of {its_full} GMRES iterations ({out['seconds']:.2f} s each) + one orthogonalisation at the mean "
                  f"basis length {(its_full + 1) // 2} ({t_ortho:.2f} s); complete-solve time {t:.1f} s from the sample's timers (bench.py: _model_time; checked against complete solves at refinement 5 in model_check)")
    cb.close()
    return {"value": n / t / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "simd_cells_per_batch": width,
            "wall_s": t, "iterations": its, "dofs": n, "model_check": model_check,
            "implementation": "oracle/csrc/smf_cpu_solver.cc: C++17 -O3, OpenMP (one thread per core), SIMD cell batches, solver loop in C++"}


def cfg2_cpu_reference(n_solves=2):
    """Complete solves of BASELINE cfg 2 (refinement 5) by the CPU arm: the like-for-like partner of `cfg2_refinement5`."""
    cb, rhs, n, threads = _cpu_problem(5)
    outs = [cb.solve(rhs, tol_rel=TOL, restart=RESTART, max_its=CHEAP_STEPS) for _ in range(1 + n_solves)][1:]
    cb.close()
    t = float(np.mean([o["seconds"] for o in outs]))
    return {"workload": workload_name(5), "dofs": n, "wall_s": t, "mdofs": n / t / 1e6, "iterations": outs[-1]["iterations"], "cores": threads,
            "split_s": outs[-1]["split"]}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores, on the configuration of the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = args.refinement
    world = args.gpus if args.gpus in REPETITIONS else 1
    res = cpu_reference_steps(R, args.steps, args.warmup)
    note = None
    if world > 1:
        note = (f"the host cores do not multiply with --gpus: the CPU arm times ONE of the {world} coarse-cell trees of this workload "
                f"(refinement {R}, {res['dofs']} DoFs) and reports its throughput")
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["wall_s"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(R, world), "dofs_per_gpu": res["dofs"], "levels": R + 1,
                       "parallelism": f"host cores: OpenMP, {res['cores']} threads, {res['simd_cells_per_batch']} cells per SIMD batch",
                       "impl": "reference"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "model_check", "implementation")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "solve_wall_s": res["wall_s"], "iterations": res["iterations"], "note": note}
    if not args.no_cfg2:
        line["cfg2_refinement5"] = cfg2_cpu_reference()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[0]) for r in rows]
        out["sm_mhz"] = float(np.median(sm))
        out["sm_max_mhz"] = float(rows[0][1])
        out["power_w_max"] = max(float(r[2]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for j, n in enumerate(names) if any(r[3 + j].strip().lower().startswith("active") for r in rows)]
        out["samples"] = len(rows)
        return out


def build_problem(R, device, rank=0, world=1):
    from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
    hier = synthetic.BoxHierarchy(3, R, None, REPETITIONS[world], rank, world)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, N_SINKERS, VISCOSITY_RATIO, AVERAGING)
    h = StokesMatrixFreeHandler(3, device=device)
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, AVERAGING, ps)
    h.build_preconditioner()
    h.parameters.update({"Linear solver tolerance": TOL, "Number of cheap Stokes solver steps": CHEAP_STEPS,
                         "GMRES solver restart length": RESTART})
    return hier, h, rhs, eta_q, ps


def timed_solves(torch, h, stream, rhs_dev, sol_dev, n_u, n_p, steps, barrier):
    """K device-resident solves from a zero guess; CUDA events on the launching stream, barrier + sync both sides."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    outs = []
    barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        sol_dev[0].zero(); sol_dev[1].zero()
        outs.append(h.solve_device(sol_dev, rhs_dev))
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    return ev0.elapsed_time(ev1) / 1e3, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refinement", type=int, default=int(os.environ.get("SMF_BENCH_REFINEMENT", "7")),
                    help="global refinement per GPU (7 = north-star single-GPU target; 8 GPUs x 7 = BASELINE cfg 5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the N-rank = 1-rank parity block")
    ap.add_argument("--parity-refinement", type=int, default=4, help="N > 1: refinement per tree of the parity block")
    ap.add_argument("--no-solver-timings", action="store_true", help="skip the Execute-solver-timings harness (incl. the IDR(2) solve)")
    ap.add_argument("--no-cfg2", action="store_true", help="skip the extra BASELINE cfg 2 (refinement 5) solve")
    ap.add_argument("--convbox-refinement", type=int, default=5, help="global refinement of the extra BASELINE cfg 4 (convection box) block")
    ap.add_argument("--shell-refinement", type=int, default=3, help="global refinement of the extra BASELINE cfg 3 (spherical shell) block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the block-GMG path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sync_t = torch.zeros(1, device="cuda")

    def barrier():
        if world > 1:
            dist.all_reduce(sync_t)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from aspect_b200 import SMF_ACTIVE
    R = args.refinement
    t_setup = time.time()
    hier, h, rhs, eta_q, ps = build_problem(R, local_rank, rank, world)
    cpu_inputs = (hier, eta_q, rhs.copy(), ps) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    del eta_q
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    n_u, n_p, n_cells = hier.active.n_u, hier.active.n_p, hier.active.n_cells
    n = n_u + n_p
    setup_s = time.time() - t_setup

    rhs_dev, sol_dev = h.initialize_dof_vector(), h.initialize_dof_vector()
    rhs_dev[0].from_host(rhs[:n_u]); rhs_dev[1].from_host(rhs[n_u:])
    if world > 1:       # per-rank assembly leaves partial sums on the subdomain interfaces: compress(add)
        h.exchange_add(rhs_dev[0], 0); h.exchange_add(rhs_dev[1], 1)
        rhs = np.concatenate([rhs_dev[0].to_host(), rhs_dev[1].to_host()])
    # global DoF count: shared DoFs counted once
    n_global = n
    if world > 1:
        act = hier.active
        n_global = 3 * int(np.prod(act.grid_u)) + int(np.prod(act.grid_p))
    # pinned host buffers of the end-to-end arm
    rhs_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    sol_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    rhs_pin.numpy()[:] = rhs
    del rhs

    with torch.cuda.stream(stream):
        # ---- warm-up, then the timed device-resident solves --------------------------------------------------
        for _ in range(max(args.warmup, 3)):
            sol_dev[0].zero(); sol_dev[1].zero()
            h.solve_device(sol_dev, rhs_dev)
        h.profile_enable(1)
        launches0, exch0 = h.kernel_launches(), h.exchange_count()
        clocks = ClockSampler(local_rank)
        t_dev, outs = timed_solves(torch, h, stream, rhs_dev, sol_dev, n_u, n_p, args.steps, barrier)
        clock_info = clocks.stop()
        launches = h.kernel_launches() - launches0
        exchanges = (h.exchange_count() - exch0) // max(args.steps, 1)
        prof = h.profile_read()
        h.profile_enable(0)
        t_dev = max_over_ranks(t_dev)
        # ---- the same through the host-facing call ----------------------------------------------------------
        sol_np, rhs_np = sol_pin.numpy(), rhs_pin.numpy()
        for _ in range(1):
            h.solve(rhs_np, None, out=sol_np)
        barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record(stream)
        for _ in range(args.steps):
            _, out_e2e = h.solve(rhs_np, None, out=sol_np)
        ev1.record(stream)
        torch.cuda.synchronize(); barrier()
        t_e2e = max_over_ranks(max(ev0.elapsed_time(ev1) / 1e3, time.time() - t0))
        # ---- bucket split of one more solve (every launch bracketed by events; not part of the timed region) --
        h.profile_enable(2)
        sol_dev[0].zero(); sol_dev[1].zero()
        h.solve_device(sol_dev, rhs_dev)
        split = h.profile_read()
        h.profile_enable(0)
        # ---- operator throughput --------------------------------------------------------------------------
        src, dst = h.initialize_dof_vector(), h.initialize_dof_vector()
        rng = np.random.default_rng(0)
        xu = rng.uniform(-1, 1, n_u); xu[hier.active.constrained_u("zero").astype(np.int64)] = 0.0
        src[0].from_host(xu); src[1].from_host(rng.uniform(-1, 1, n_p))
        del xu

        def time_op(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return max_over_ranks(a.elapsed_time(b) / 1e3 / reps)

        fp64_peak = h.measure_fp64_peak()
        t_stokes = time_op(lambda: h.stokes_matrix.vmult(dst, src))
        t_A = time_op(lambda: h.A_block_matrix.vmult(dst[0], src[0]))
        t_M = time_op(lambda: h.Schur_complement_block_matrix.vmult(dst[1], src[1]))
        h.profile_enable(1)
        for _ in range(20):
            h.stokes_matrix.vmult(dst, src)
        pk = h.profile_read()
        h.profile_enable(0)
        t_stokes_kernel = pk["stokes_vmult"][0] / 1e3 / max(pk["stokes_vmult"][1], 1)
        # host-buffer Stokes vmult (smf_vmult_stokes_host): PCIe-bound, reported for completeness
        dst_pin = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
        from aspect_b200._lib import ptr
        def host_vmult():
            h._check(h.lib.smf_vmult_stokes_host(h.ctx, ptr(dst_pin[:n_u]), ptr(dst_pin[n_u:]), ptr(rhs_np[:n_u]), ptr(rhs_np[n_u:])))
        host_vmult()
        t0 = time.time()
        for _ in range(3):
            host_vmult()
        t_stokes_host = max_over_ranks((time.time() - t0) / 3)
        # ---- the reference's own "Execute solver timings" harness (local_smoothing.cc:1119-1260), 3 timings each --------
        solver_timings = None
        if not args.no_solver_timings:
            solver_timings = h.execute_solver_timings(rhs_np, n_timings=3, out=lambda *_: None)
            solver_timings = {k: max_over_ranks(v) for k, v in sorted(solver_timings.items())}

    lattice_active, lattice_repl, repl_A = h.lattice_schur_active(), h.lattice_replicated_levels(), h.replicated_A_levels()
    its = outs[-1]["iterations"]
    wall = t_dev / args.steps
    wall_e2e = t_e2e / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    # algorithmic bytes: A-block = 16 N_u + 8 N_cell n_eta (level tables: 1 value per cell)
    bytes_A = 16.0 * n_u + 8.0 * n_cells
    bytes_stokes = 16.0 * n + 8.0 * n_cells
    a_ms, a_cnt = prof["A_fine"]
    t_A_kernel = a_ms / 1e3 / max(a_cnt, 1)
    s_ms, s_cnt = prof["stokes_vmult"]                  # Stokes vmult kernel, timed live inside the timed solves as well
    t_S_live = s_ms / 1e3 / max(s_cnt, 1)
    # FP64 work per launch: 2,268 DFMA-class instructions per cell for the A block, + 369 for the
    # pressure interpolation / divergence rows of the Stokes operator; 2 flops each
    flops_A, flops_stokes = 2.0 * 2268 * n_cells, 2.0 * (2268 + 369) * n_cells
    traffic = traffic_stokes = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic, traffic_stokes = tj.get(f"A_fine_R{R}"), tj.get(f"stokes_R{R}")
    except (OSError, ValueError):
        pass
    line = {
        "metric": METRIC, "value": n_global / wall / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(R, world), "dofs_global": n_global, "dofs_per_gpu": n, "cells_per_gpu": n_cells, "levels": R + 1,
                   "parallelism": "1 subdomain per GPU" if world == 1 else f"domain decomposition, {world} subdomains (one coarse-cell tree per GPU), halo exchange = one fused pack/push/wait/accumulate kernel over NVLink peer memory (CUDA IPC), dot products by ncclAllReduce",
                   "l2": "vectors (%.0f MB each) exceed the 126 MB L2" % (8e-6 * n) if 8 * n > 126e6 else "vectors fit L2: each solve streams >100 distinct vectors between reuses",
                   "impl": "b200"},
        "solve_wall_s": wall, "iterations": its,
        "n_A_preconditioner": outs[-1]["n_A_preconditioner"], "n_S_preconditioner": outs[-1]["n_S_preconditioner"],
        "solve_setup_s": outs[-1]["setup_seconds"], "final_residual_over_tol": outs[-1]["final_residual"] / outs[-1]["tolerance"],
        "e2e": {"value": n_global / wall_e2e / 1e6, "unit": UNIT, "wall_s": wall_e2e,
                "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 8 * n,
                "call": "StokesMatrixFreeHandler.solve -> smf_solve_host (pinned host rhs + guess in, solution out)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "velocity_sb_kernel<0> (A block, finest level)", "achieved": bytes_A / t_A_kernel / 1e9,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": bytes_A / t_A_kernel / 1e9 / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": bytes_A, "launches_timed": a_cnt,
                     "avg_launch_us": t_A_kernel * 1e6, "share_of_solve": a_ms / 1e3 / t_dev,
                     "fp64": {"achieved_tflops": flops_A / t_A_kernel / 1e12, "peak_tflops": fp64_peak, "frac": flops_A / t_A_kernel / 1e12 / fp64_peak,
                              "peak_source": "measured in this run (smf_measure_fp64_peak: DFMA micro-kernel, best of 5 launches)"}},
        "roofline_stokes_vmult": {"bound": "hbm", "kernel": "velocity_sb_kernel<1> (Stokes operator = the op the metric names, active level)",
                                  "achieved": bytes_stokes / t_S_live / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_stokes / t_S_live / 1e9 / peak,
                                  "traffic": traffic_stokes, "algorithmic_bytes_per_launch": bytes_stokes, "launches_timed": s_cnt,
                                  "avg_launch_us": t_S_live * 1e6, "share_of_solve": s_ms / 1e3 / t_dev, "gdofs": n / t_S_live / 1e9,
                                  "fp64": {"achieved_tflops": flops_stokes / t_S_live / 1e12, "peak_tflops": fp64_peak,
                                           "frac": flops_stokes / t_S_live / 1e12 / fp64_peak}},
        "vmult": {"stokes_gdofs": n_global / t_stokes / 1e9, "stokes_kernel_gdofs": n_global / t_stokes_kernel / 1e9,
                  "stokes_roofline_frac": bytes_stokes / t_stokes_kernel / 1e9 / peak,
                  "A_gdofs": n_global / n * n_u / t_A / 1e9, "mass_gdofs": n_global / n * n_p / t_M / 1e9,
                  "stokes_host_buffers_gdofs": n_global / t_stokes_host / 1e9,
                  "note": "stokes_gdofs includes zeroing dst (StokesOperator::vmult); *_kernel_* is the cell kernel alone"},
        "time_split_ms": {k: round(v[0], 3) for k, v in split.items()},
        "launch_split": {k: v[1] for k, v in split.items()},
        "clocks": clock_info, "host_setup_s": setup_s, "solver_timings_s": solver_timings, "halo_exchanges_per_solve": int(exchanges), "halo_transport": h.exchange_transport(),
        "schur_vcycle": {"path": "lattice (node-centred mass stencil + fused Chebyshev update, one CUDA graph)" if lattice_active else "cell batches",
                         "replicated_coarse_levels": lattice_repl,
                         "ms": None if not solver_timings else solver_timings.get("S_preconditioner", 0.0) * 1e3,
                         # algorithmic bytes of one cycle: 32 smoothing steps per level, 40 B per pressure DoF and step (x, xo in/out, b,
                         # one coefficient), levels sum to 8/7 of the finest
                         "algorithmic_bytes": 32 * 40.0 * n_p * 8.0 / 7.0,
                         "hbm_frac": None if not solver_timings or not solver_timings.get("S_preconditioner") else
                         32 * 40.0 * n_p * 8.0 / 7.0 / solver_timings["S_preconditioner"] / 1e9 / peak},
    }
    h.close()
    if world > 1 and lattice_repl:
        line["config"]["parallelism"] += (f"; Schur V-cycle: the {lattice_repl} coarsest levels replicated on every rank (global lattice, one all-reduce per cycle)"
                                          f"; prec_A: the {repl_A} coarsest levels replicated on every rank (global mesh in a child context, one all-reduce per cycle)")
        line["replicated_coarse_levels"] = {"schur": lattice_repl, "A": repl_A}
    # ---- BASELINE cfg 2 (refinement 5) on the same GPU, for the apples-to-apples comparison with the CPU arm ------
    if not args.no_cfg2 and R != 5 and world == 1:
        hier5, h5, rhs5, _, _ = build_problem(5, local_rank)
        n5 = hier5.active.n_u + hier5.active.n_p
        pin5 = torch.empty(n5, dtype=torch.float64).pin_memory().numpy()
        rp5 = torch.empty(n5, dtype=torch.float64).pin_memory().numpy(); rp5[:] = rhs5
        for _ in range(2):
            _, o5 = h5.solve(rp5, None, out=pin5)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(3):
            _, o5 = h5.solve(rp5, None, out=pin5)
        t5 = (time.time() - t0) / 3
        line["cfg2_refinement5"] = {"workload": workload_name(5), "dofs": n5, "e2e_wall_s": t5, "e2e_mdofs": n5 / t5 / 1e6,
                                    "iterations": o5["iterations"]}
        h5.close()
        # ---- BASELINE cfg 1: benchmarks/solcx, 2D Q2-Q1, global refinement 6 (general-batch 2D kernels) ----------------
        from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic
        hier1 = synthetic.BoxHierarchy(2, 6)
        avg1 = "harmonic average only viscosity"
        eta1, rhs1, ps1 = benchmarks.solcx_problem(hier1, averaging=avg1)
        h1 = StokesMatrixFreeHandler(2, device=local_rank)
        h1.setup_dofs(hier1, "tangential")
        h1.assemble(eta1, avg1, ps1)
        h1.build_preconditioner()
        h1.parameters.update({"Linear solver tolerance": 1e-7, "Pressure normalization": "surface"})
        for _ in range(2):
            _, o1 = h1.solve(rhs1)
        t0 = time.time()
        for _ in range(3):
            _, o1 = h1.solve(rhs1)
        t1 = (time.time() - t0) / 3
        n1 = hier1.active.n_u + hier1.active.n_p
        line["cfg1_solcx_refinement6"] = {"workload": "SolCx 2D unit square, viscosity jump 1e6, free slip, Q2-Q1, block GMG (GMRES(100), tol 1e-07), global refinement 6",
                                          "dofs": n1, "e2e_wall_s": t1, "e2e_mdofs": n1 / t1 / 1e6, "iterations": o1["iterations"],
                                          "pressure_normalization_adjustment": o1["pressure_normalization_adjustment"]}
        h1.close()
        # ---- BASELINE cfg 3: benchmarks/nsinker_spherical_shell (curved cells, MappingQ(4) geometry streamed per point) ------
        try:
            from aspect_b200 import shell as pshell
            R3 = args.shell_refinement
            hier3 = pshell.ShellHierarchy(R3)
            eta3, force3, ps3 = benchmarks.nsinker_shell_problem(hier3)
            act3 = hier3.active
            h3 = StokesMatrixFreeHandler(3, device=local_rank)
            h3.setup_dofs(hier3, "zero")
            h3.assemble(eta3, "harmonic average", ps3)
            h3.build_preconditioner()
            h3.parameters.update({"Linear solver tolerance": 1e-6, "Number of cheap Stokes solver steps": 2000,
                                  "Maximum number of expensive Stokes solver steps": 0, "GMRES solver restart length": 200})
            r3 = h3.initialize_dof_vector(0)
            h3.assemble_stokes_rhs(force3, r3)
            rhs3 = np.concatenate([r3.to_host(), np.zeros(act3.n_p)])
            for _ in range(2):
                _, o3 = h3.solve(rhs3)
            t0 = time.time()
            for _ in range(3):
                _, o3 = h3.solve(rhs3)
            t3 = (time.time() - t0) / 3
            s3, d3 = h3.initialize_dof_vector(), h3.initialize_dof_vector()
            s3[0].from_host(np.sin(np.arange(act3.n_u))); s3[1].from_host(np.sin(np.arange(act3.n_p)))

            def time3(fn, reps=20):
                for _ in range(3):
                    fn()
                h3.synchronize()
                a = time.perf_counter()
                for _ in range(reps):
                    fn()
                h3.synchronize()
                return (time.perf_counter() - a) / reps
            ts3 = time3(lambda: h3.stokes_matrix.vmult(d3, s3))
            ta3 = time3(lambda: h3.A_block_matrix.vmult(d3[0], s3[0]))
            n3 = act3.n_u + act3.n_p
            alg3 = 16 * act3.n_u + 8 * act3.n_cells * (1 + 270)        # vectors + viscosity + G (27 points x (9 + 1) doubles)
            line["cfg3_spherical_shell"] = {
                "workload": "nsinker_spherical_shell 3D (96 coarse cells, R0 0.54, R1 1, MappingQ(4) geometry), 5 sinkers, viscosity ratio 1e2, "
                            "transition-zone jump 100, harmonic average, Q2-Q1, block GMG (GMRES(200), tol 1e-06), global refinement %d" % R3,
                "dofs": n3, "cells": act3.n_cells, "e2e_wall_s": t3, "e2e_mdofs": n3 / t3 / 1e6, "iterations": o3["iterations"],
                "stokes_gdofs": n3 / ts3 / 1e9, "A_gdofs": act3.n_u / ta3 / 1e9,
                "A_roofline_frac": alg3 / ta3 / 1e9 / peak, "A_algorithmic_bytes": alg3, "A_ms": ta3 * 1e3,
                "viscosity_range": [h3.minimum_viscosity, h3.maximum_viscosity]}
            h3.close()
        except Exception as exc:            # the headline line above must survive a failure of the extra configuration
            line["cfg3_spherical_shell"] = {"error": repr(exc)}
        # ---- BASELINE cfg 4 shape: cookbooks/convection_box_3d, the Stokes solve of successive time steps (same mesh and set-up,
        #      changing right-hand side, warm start from the previous step); the temperature advection is not on this path ----
        try:
            R4 = args.convbox_refinement
            hier4 = synthetic.BoxHierarchy(3, R4)
            eta4, rhs4, ps4 = benchmarks.convection_box_problem(hier4)
            h4 = StokesMatrixFreeHandler(3, device=local_rank)
            h4.setup_dofs(hier4, "tangential")
            h4.assemble(eta4, "none", ps4)
            h4.build_preconditioner()
            h4.parameters.update({"Linear solver tolerance": 1e-7, "Number of cheap Stokes solver steps": 200,
                                  "Maximum number of expensive Stokes solver steps": 0})
            steps4, sol4 = [], None
            for k in range(4):
                _, rhs4, _ = benchmarks.convection_box_problem(hier4, perturbation=0.01 * (1.0 + 0.5 * k))
                t0 = time.time()
                sol4, o4 = h4.solve(rhs4, sol4)
                steps4.append({"iterations": o4["iterations"], "wall_s": time.time() - t0, "converged": bool(o4["converged"])})
            n4 = hier4.active.n_u + hier4.active.n_p
            line["cfg4_convection_box_3d_stokes_steps"] = {
                "workload": "cookbooks/convection_box_3d shape: unit box, viscosity 1, tangential velocity on all faces, Ra 1e6 buoyancy, "
                            "Q2-Q1, block GMG (tol 1e-07), global refinement %d, 4 successive right-hand sides, warm start" % R4,
                "dofs": n4, "steps": steps4}
            h4.close()
        except Exception as exc:
            line["cfg4_convection_box_3d_stokes_steps"] = {"error": repr(exc)}
        # ---- BASELINE cfg 4 as shipped: 3 global + 3 ADAPTIVE refinement levels (convection_box_3d.prm:119-121).  The Kelly
        #      indicator of the temperature field is outside this path; the synthetic marker refines the two thermal boundary
        #      layers level by level, which is where that indicator puts the cells.  Hanging nodes, refinement edges ----
        try:
            from aspect_b200 import adaptive as padapt
            t0 = time.time()
            hier4a = padapt.AdaptiveBoxHierarchy.refined(
                3, 3, lambda x, l: min(x[2], 1.0 - x[2]) < 0.25 / (1 << (l - 3)), 3)
            t_mesh = time.time() - t0
            eta4a, rhs4a, ps4a = benchmarks.convection_box_problem(hier4a)
            h4a = StokesMatrixFreeHandler(3, device=local_rank)
            h4a.setup_dofs(hier4a, "tangential")
            h4a.assemble(eta4a, "none", ps4a)
            h4a.build_preconditioner()
            h4a.parameters.update({"Linear solver tolerance": 1e-7, "Number of cheap Stokes solver steps": 200,
                                   "Maximum number of expensive Stokes solver steps": 0})
            h4a.condense_rhs(rhs4a)
            rhs4a[hier4a.active.constrained_u("tangential").astype(np.int64)] = 0.0
            _, o4a = h4a.solve(rhs4a)
            t0 = time.time()
            _, o4a = h4a.solve(rhs4a)
            act4a = hier4a.active
            line["cfg4_convection_box_3d_adaptive"] = {
                "workload": "cookbooks/convection_box_3d shape on an ADAPTIVE mesh: 3 global + 3 adaptive refinement levels (boundary layers), "
                            "hanging-node constraints, local smoothing with refinement edges, tangential velocity, block GMG (tol 1e-07)",
                "active_cells": int(act4a.n_cells), "dofs": int(act4a.n_u + act4a.n_p), "levels": len(hier4a.levels),
                "cells_per_level": [int(L.n_cells) for L in hier4a.levels], "hanging_velocity_nodes": len(act4a.hanging_u),
                "iterations": o4a["iterations"], "converged": bool(o4a["converged"]), "e2e_wall_s": time.time() - t0, "host_mesh_s": t_mesh}
            h4a.close()
        except Exception as exc:
            line["cfg4_convection_box_3d_adaptive"] = {"error": repr(exc)}
    # ---- N > 1: the N-rank run against ONE rank on the same global mesh (tools/dist_check.py: operator 1e-12, iterations +-1,
    #      solution 1e-8, both V-cycles, copies of shared DoFs bitwise equal), at a size rank 0 solves alone in seconds; outside
    #      the timed region, every rank takes part (collective) ----
    if world > 1 and not args.no_parity:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import dist_check
            line["multi_gpu_parity"] = dist_check.parity(args.parity_refinement, "gmres")
        except Exception as exc:            # must be collective-safe: every rank fails or none (same code, same inputs)
            line["multi_gpu_parity"] = {"ok": False, "error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_steps(R, 2, 1, *cpu_inputs)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "model_check", "implementation")}
        if "cfg2_refinement5" in line:
            line["cfg2_refinement5"]["cpu_reference"] = cfg2_cpu_reference()
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
