#!/usr/bin/env python
"""bench.py -- iLQR OCP solves/s (batched single-track, fp64) on N B200s of one node.

Workload (BASELINE.json configs[2], SURVEY 8d "config 3"): 65,536 independent single-track
lane-following OCPs (n=4, m=2, T=80, dt=0.1, bounds as examples/single_track_ocp.cpp:105-109),
x0 = (0, Y, psi, v) from std::mt19937_64(20240607), U_init = 0, iLQR params 10 / 1e-5 / max_ms=inf.
A "step" is one solve of the whole batch.  With N > 1 every rank solves its own 65,536-problem shard
(independent problems, no data-path collective): weak scaling; --scaling strong splits one 65,536
batch across the ranks instead.  Throughput is measured with --depth (default 6) independent solves in
flight per GPU -- each a whole step on its own stream, driven by its own host thread, starts staggered --
because the last iterations of a solve are bound by the latency of T sequential time steps and leave
the GPU nearly idle; the time of one solve alone is reported as config.single_solve_ms.

  value : solves/s, inputs (x0) resident in HBM in the engine's layout when the timed region starts
  e2e   : same metric through the C ABI with HOST buffers: x0 host->device, solve, and X, U, cost,
          iterations, status device->host inside the timed region, every step
  roofline : the dominant kernel, timed per launch with CUDA events inside the engine
  cpu_baseline : the oracle (CPU restatement of the reference, OpenMP over problems) on a bounded sample

--impl reference times the reference's own CPU path instead: the reference cannot be built here
(needs Eigen 3.4, absent, no network), so this is the oracle port on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "iLQR OCP solves/sec (batched single-track, fp64)"
UNIT = "solves/s"
PROBLEMS = 65536
T, NX, NU = 80, 4, 2
MAX_ITER, TOL = 10, 1e-5
# algorithmic HBM bytes per problem-iteration: backward reads X,U and writes K,k; forward
# reads X,U,K,k once and writes the accepted X,U once
BWD_BYTES = 8 * ((T + 1) * NX + T * NU + T * NU * NX + T * NU)
FWD_BYTES = 8 * ((T + 1) * NX + T * NU + T * NU * NX + T * NU) + 8 * ((T + 1) * NX + T * NU)
# algorithmic flops per time step, SURVEY 8d convention (MAC = 2, one trig/div-heavy call = 32)
BWD_FLOPS_STEP, FWD_FLOPS_STEP = 1764.0, 494.0 + 12.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches of one
    solve, from the committed ncu capture of this same workload (profiles/r01_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    # the engine's "forward" timing bucket is the line search whichever kernel ran it: the warp-cooperative kernel
    # (large active sets) and the lane kernel (small ones); average over the launches of both
    names = ["forward_coop_kernel", "forward_kernel"] if kernel == "forward_kernel" else [kernel]
    try:
        with open(path) as f:
            t = json.load(f)
        launches = sum(t[n]["launches"] for n in names if n in t)
        return sum(t[n]["launches"] * t[n]["dram_bytes_per_launch"] for n in names if n in t) / launches
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(n_problems: int, steps: int, warmup: int, threads: int = 0):
    """Times the oracle (the reference's CPU algorithm, OpenMP parallel-for over problems) on the
    first n_problems problems of the workload.  Returns (solves/s, threads, ms/step)."""
    from oracle import oracle_py as o
    import multi_agent_solver_b200 as mas

    x0 = mas.synthetic_single_track_x0(PROBLEMS)[:n_problems]
    # all host cores this process may use (torchrun sets OMP_NUM_THREADS=1; the num_threads clause overrides it)
    threads = threads or len(os.sched_getaffinity(0))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.ilqr_solve_batch(o.MODEL_ST_LANE, x0, max_iterations=MAX_ITER, tolerance=TOL, trig=o.TRIG_GLIBC, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    return n_problems / mean, threads, mean * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_sample
    value, threads, ms = cpu_reference_run(n, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched single-track iLQR, 65,536 independent OCPs (BASELINE configs[2]); each step solves a bounded sample",
                   "problems_per_step": n, "horizon": T, "max_iterations": MAX_ITER, "tolerance": TOL},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"first {n} of the 65,536 problems per step; oracle/ (C++ restatement of the reference, glibc libm, "
                                   f"-O3 -ffp-contract=off, OpenMP static over problems); the reference itself needs Eigen 3.4, absent here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)




class Lane:
    """One solve pipeline: its own CUDA stream, engine context, resident batch and pinned result buffers."""

    def __init__(self, torch, mas, device, desc, per_rank, args, with_host_buffers):
        self.stream = torch.cuda.Stream()
        self.ctx = mas.Context(device, self.stream.cuda_stream)
        self.batch = mas.Batch(self.ctx, desc, per_rank)
        if args.lanes or args.chains:
            self.batch.set_tuning(args.lanes, args.chains)
        if args.ls_mode:
            self.batch.set_line_search_mode(args.ls_mode)
        self.out = None
        if with_host_buffers:
            self.out = dict(X=torch.empty((per_rank, T + 1, NX), dtype=torch.float64).pin_memory().numpy(),
                            U=torch.empty((per_rank, T, NU), dtype=torch.float64).pin_memory().numpy(),
                            cost=torch.empty(per_rank, dtype=torch.float64).pin_memory().numpy(),
                            iterations=torch.empty(per_rank, dtype=torch.int32).pin_memory().numpy(),
                            status=torch.empty(per_rank, dtype=torch.int32).pin_memory().numpy())


def run_b200(args):
    import torch
    import multi_agent_solver_b200 as mas

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    per_rank = PROBLEMS if args.scaling == "weak" else PROBLEMS // world
    total = per_rank * world

    x0_all = mas.synthetic_single_track_x0(PROBLEMS)
    if args.scaling == "weak":
        # every rank gets the 65,536 problems, rotated so the ranks do not solve identical shards in the same order
        x0 = np.roll(x0_all, -rank * 4099, axis=0).copy()
    else:
        x0 = x0_all[rank * per_rank:(rank + 1) * per_rank].copy()
    x0_host = torch.from_numpy(x0).pin_memory().numpy()

    desc = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
    prm = mas.IlqrParams.make(MAX_ITER, TOL)
    depth = args.depth if args.depth > 0 else (3 if args.steps < 8 else 4)
    lanes = [Lane(torch, mas, local_rank, desc, per_rank, args, with_host_buffers=not args.resident_only) for _ in range(depth)]
    for ln in lanes:
        ln.batch.set_initial_states(x0_host)  # resident input of the `value` measurement

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(ln):
        ln.batch.set_controls(None)  # U_init = 0 (device memset); x0 is already resident
        ln.batch.solve(prm)

    def e2e_step(ln):
        ln.batch.set_initial_states(x0_host)  # H2D from pinned memory
        ln.batch.set_controls(None)
        ln.batch.solve(prm)
        # D2H of X, U, cost, iterations, status into pinned memory: staged in HBM on the solve stream, copied on the
        # batch's copy stream while this lane's next solve starts; e2e_finish() waits for the last one inside the
        # timed region, and every begin waits for the previous download of the lane
        ln.batch.begin_get_solution(ln.out)

    def e2e_finish(ln):
        ln.batch.wait_solution()

    def timed(fn, steps, use_lanes, stagger_ms, finish=None):
        """Runs exactly `steps` steps spread over `use_lanes` pipelines, each driven by its own host thread and
        started stagger_ms/len(use_lanes) apart so one solve's latency-bound tail overlaps another's bulk.
        Device time = latest end event - earliest start event over the lanes' streams (then max over ranks)."""
        n = len(use_lanes)
        counts = [steps // n + (1 if i < steps % n else 0) for i in range(n)]
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        go = threading.Barrier(n + 1)
        errors = []

        def work(i):
            try:
                torch.cuda.set_device(local_rank)  # the current device is per host thread
                ln = use_lanes[i]
                go.wait()
                if i:
                    time.sleep(i * stagger_ms * args.stagger * 1e-3 / n)
                e0[i].record(ln.stream)
                for _ in range(counts[i]):
                    fn(ln)
                if finish is not None:
                    finish(ln)  # host-blocking: the lane's last results are in host memory before the end event
                e1[i].record(ln.stream)
                e1[i].synchronize()
            except Exception as exc:  # surfaced after join
                errors.append(exc)

        threads = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        for t in threads:
            t.start()
        barrier()
        w0 = time.perf_counter()
        go.wait()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        if errors:
            raise errors[0]
        active = [i for i in range(n) if counts[i] > 0]
        starts = [e0[active[0]].elapsed_time(e0[i]) for i in active]
        ends = [e0[active[0]].elapsed_time(e1[i]) for i in active]
        ms = max(ends) - min(starts)
        barrier()
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms / steps, wall / steps

    # ---- one solve at a time: latency of a 65,536-problem solve, and the reference for kernel shares ------------
    for _ in range(max(args.warmup, 3)):
        for ln in lanes:
            resident_step(ln)
    barrier()
    single_ms, _ = timed(resident_step, max(3, min(args.steps, 6)), lanes[:1], 0.0)

    # ---- resident throughput: `depth` solves in flight ----------------------------------------------------------
    launches0 = sum(ln.batch.stats()["kernel_launches"] for ln in lanes)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, wall_step = timed(resident_step, args.steps, lanes, single_ms)
    launches = sum(ln.batch.stats()["kernel_launches"] for ln in lanes) - launches0
    st = lanes[0].batch.stats()
    value = total / (ms_step * 1e-3)

    if args.resident_only:  # short form for ncu captures: only the resident steps above
        if rank == 0:
            sampler.stop()
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_step, "single_solve_ms": single_ms,
                              "gpu_launches": int(launches), "note": "resident-only run (profiling aid), not a bench line"}), flush=True)
        for ln in lanes:
            ln.batch.close()
        return

    # ---- e2e: host buffers in, host buffers out, every step ---------------------------------------------------------
    for ln in lanes:
        e2e_step(ln)
        e2e_finish(ln)
    e2e_ms, e2e_wall = timed(e2e_step, args.steps, lanes, single_ms, finish=e2e_finish)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = total / (max(e2e_ms * 1e-3, e2e_wall))
    h2d = per_rank * NX * 8
    d2h = per_rank * (((T + 1) * NX + T * NU + 1) * 8 + 2 * 4)

    # the same loop for a caller that only wants the controls (best_controls, best_cost, iterations, status; X = NULL):
    # a third of the download, reported next to the headline e2e because at N > 1 the host's D2H bandwidth sets e2e
    def e2e_step_u(ln):
        ln.batch.set_initial_states(x0_host)
        ln.batch.set_controls(None)
        ln.batch.solve(prm)
        ln.batch.begin_get_solution({k: v for k, v in ln.out.items() if k != "X"})

    for ln in lanes:
        e2e_step_u(ln)
        e2e_finish(ln)
    e2e_u_ms, e2e_u_wall = timed(e2e_step_u, args.steps, lanes, single_ms, finish=e2e_finish)
    e2e_u_value = total / (max(e2e_u_ms * 1e-3, e2e_u_wall))
    d2h_u = per_rank * ((T * NU + 1) * 8 + 2 * 4)

    # ---- per-kernel timing for the roofline: one solve at a time, CUDA events around every launch inside the engine
    batch = lanes[0].batch
    batch.set_profiling(True)
    for _ in range(max(2, min(args.steps, 5))):
        resident_step(lanes[0])
    prof = batch.profile()
    batch.set_profiling(False)
    fp64_peak = lanes[0].ctx.probe_fp64_peak()
    out = batch.get_solution()

    if rank == 0:
        hbm_peak, peak_src = load_peaks()
        iters_total = int(out["iterations"].sum())
        n_solves = max(prof["solves"], 1)
        pit = prof["problem_iterations"] / n_solves  # problem-iterations per solve of this rank's shard
        kernels = {
            "forward_kernel": {"ms_per_solve": prof["forward_ms"] / n_solves, "launches_per_solve": prof["forward_launches"] / n_solves,
                               "alg_bytes_per_solve": pit * FWD_BYTES},
            "backward_kernel": {"ms_per_solve": prof["backward_ms"] / n_solves, "launches_per_solve": prof["backward_launches"] / n_solves,
                                "alg_bytes_per_solve": pit * BWD_BYTES},
            "prologue_kernel": {"ms_per_solve": prof["prologue_ms"] / n_solves, "launches_per_solve": 1,
                                "alg_bytes_per_solve": per_rank * 8 * ((T + 1) * NX + T * NU)},
        }
        dom = max(("forward_kernel", "backward_kernel"), key=lambda k: kernels[k]["ms_per_solve"])
        k = kernels[dom]
        n_launch = max(k["launches_per_solve"], 1)
        achieved = (k["alg_bytes_per_solve"] / n_launch) / (k["ms_per_solve"] / n_launch * 1e-3) / 1e9
        alg_flops = T * (st["iterations"] * BWD_FLOPS_STEP + (st["alpha_trials"] + per_rank) * FWD_FLOPS_STEP)
        roofline = {
            "bound": "hbm", "kernel": "line search (forward_coop_kernel + forward_kernel)" if dom == "forward_kernel" else dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": load_traffic(dom), "peak_source": peak_src,
            "avg_launch_ms": k["ms_per_solve"] / n_launch, "alg_bytes_per_launch": k["alg_bytes_per_solve"] / n_launch,
            "kernel_share_of_single_solve": {name: v["ms_per_solve"] / single_ms for name, v in kernels.items()},
            "note": "kernel durations: CUDA events around every launch on the engine's stream, one solve in flight; the kernels are bound by the "
                    "fp64 pipe and by the latency of T sequential steps, not by HBM (profiles/README.md)",
            "fp64": {"alg_tflops": alg_flops / (ms_step * 1e-3) / 1e12, "dfma_peak_tflops_measured": fp64_peak,
                     "frac": alg_flops / (ms_step * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                     "note": "algorithmic flops of the sequential reference (SURVEY 8d convention) / step time"},
        }
        cpu_baseline = None
        if n_gpus == 1:  # reported on rank 0 at N=1 only
            cpu_val, cpu_threads, cpu_ms = cpu_reference_run(args.cpu_sample, 1, 1)
            cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                            "sample": f"first {args.cpu_sample} of the 65,536 problems, one pass ({cpu_ms:.0f} ms); oracle/ C++ restatement of the "
                                      "reference (glibc libm, OpenMP static over problems); the reference needs Eigen 3.4, absent here"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched single-track iLQR, 65,536 independent OCPs with randomised initial states (BASELINE configs[2])",
                       "problems_per_gpu": per_rank, "problems_total": total, "horizon": T, "state_dim": NX, "control_dim": NU,
                       "max_iterations": MAX_ITER, "tolerance": TOL, "max_ms": "inf", "parallelism": f"independent shards x{n_gpus}",
                       "solves_in_flight": depth,
                       "pipelining": f"{depth} independent 65,536-problem solves in flight per GPU, each a whole step on its own stream and host "
                                     "thread, starts staggered; ms_per_step = device time of the K steps / K",
                       "single_solve_ms": single_ms,
                       "l2": "working set 670 MB per solve (X,U,K,k) > 126 MB L2, no flush needed",
                       "forward_lanes": st["forward_lanes"], "forward_chains": st["forward_chains"],
                       "mean_iterations": iters_total / per_rank,
                       "problem_iterations_per_s": value * iters_total / per_rank},
            "e2e_controls_only": {"value": e2e_u_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_u, "ms_per_step": e2e_u_ms,
                                  "note": "same loop without downloading the state trajectories (X = NULL in the C ABI); not the headline"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "wall_ms_per_step": e2e_wall * 1e3},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "wall_ms_per_step": wall_step * 1e3,
        }
        print(json.dumps(line), flush=True)
    for ln in lanes:
        ln.batch.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48, help="timed steps (one step = one solve of the 65,536-problem batch, 6-7 ms)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--depth", type=int, default=0, help="independent solves in flight per GPU (1 = one at a time; 0 = by step count: "
                    "3 below 8 steps, else 4 -- fewer pipelines fill and drain faster when K is small)")
    ap.add_argument("--stagger", type=float, default=1.0, help="start offset between pipelines, in units of single_solve_ms / depth")
    ap.add_argument("--cpu-sample", type=int, default=8192, help="problems per CPU-baseline pass")
    ap.add_argument("--resident-only", action="store_true", help="run only warm-up + timed resident steps (for ncu)")
    ap.add_argument("--ls-mode", type=int, default=0, help="line search scheduling: 0 auto, 1 concurrent lanes, 2 compacted rounds")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--chains", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
