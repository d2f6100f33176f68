#!/usr/bin/env python
"""bench.py -- iLQR OCP solves/s (batched single-track, fp64) on N B200s of one node.

Workload (BASELINE.json configs[2], SURVEY 8d "config 3"): 65,536 independent single-track
lane-following OCPs (n=4, m=2, T=80, dt=0.1, bounds as examples/single_track_ocp.cpp:105-109),
x0 = (0, Y, psi, v) from std::mt19937_64(20240607), U_init = 0, iLQR params 10 / 1e-5 / max_ms=inf.
A "step" is one solve of the whole batch.  With N > 1 the 65,536 problems are SHARDED over the ranks
(contiguous ranges, independent problems, no data-path collective): strong scaling, as BASELINE names
it; one full 65,536-problem shard per rank (weak scaling) is measured too and reported under "weak".

  value    : solves/s, inputs (x0) resident in HBM in the engine's layout when the timed region starts
  e2e      : same metric through the C ABI with HOST buffers: x0 host->device, solve, and X, U, cost,
             iterations, status device->host inside the timed region, every step
  parity   : the results of the timed batch against the CPU checkers, outside the timed region:
             (i) bit for bit against the reference's own sources compiled with portable trig
             (oracle/_ref, else the restated oracle), all problems of rank 0's shard;
             (ii) against the reference in its own libm mode, next to the reference's own 1-ulp band
  roofline : the dominant kernel, timed per launch with CUDA events inside the engine
  cpu_baseline : the reference's CPU implementation (oracle/_ref = its unmodified sources; else the
             restated oracle) on a bounded sample, all host threads
  configs  : short runs of BASELINE configs 1, 2, 4 (with the library's per-round NCCL all-gather at N > 1), 5

--impl reference times the reference's own CPU path instead (oracle/_ref/libref.so, built from
/root/reference's unmodified headers and examples against oracle/eigen_shim); the product library is
never loaded on that path.
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
# Hardware work queues of the CUDA context (default 8): every pipeline below is a stream, and streams that share a queue
# serialise against each other.  With 8-24 pipelines in flight the default costs 9-17 % (profiles/r02_strong_scaling_tuning.txt).
# Must be set before the CUDA context exists; a caller's own setting wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)

METRIC = "iLQR OCP solves/sec (batched single-track, fp64)"
UNIT = "solves/s"
PROBLEMS = 65536
T, NX, NU = 80, 4, 2
MAX_ITER, TOL = 10, 1e-5
WORKLOAD = "batched single-track iLQR, 65,536 independent OCPs with randomised initial states (BASELINE configs[2])"
# algorithmic HBM bytes per problem-iteration: backward reads X,U and writes K,k; forward
# reads X,U,K,k once and writes the accepted X,U once
BWD_BYTES = 8 * ((T + 1) * NX + T * NU + T * NU * NX + T * NU)
FWD_BYTES = 8 * ((T + 1) * NX + T * NU + T * NU * NX + T * NU) + 8 * ((T + 1) * NX + T * NU)
# algorithmic flops per time step, SURVEY 8d convention (MAC = 2, one trig/div-heavy call = 32)
BWD_FLOPS_STEP, FWD_FLOPS_STEP = 1764.0, 494.0 + 12.0


def shared_config(n_gpus: int) -> dict:
    """The part of `config` that both arms print identically (what is measured, not how)."""
    return {"workload": WORKLOAD, "problems_total": PROBLEMS, "horizon": T, "state_dim": NX, "control_dim": NU, "max_iterations": MAX_ITER,
            "tolerance": TOL, "max_ms": "inf", "n_gpus": n_gpus}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches of one
    solve, from the committed ncu capture of this same workload (profiles/r02_traffic.json, else r01)."""
    names = ["forward_coop_kernel", "forward_kernel"] if kernel == "forward_kernel" else [kernel]
    for fn in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                t = json.load(f)
            launches = sum(t[n]["launches"] for n in names if n in t)
            if launches:
                return sum(t[n]["launches"] * t[n]["dram_bytes_per_launch"] for n in names if n in t) / launches, fn
        except Exception:
            continue
    return None, None


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


# ---- CPU arm: the reference's own implementation --------------------------------------------------------------
def cpu_checker():
    """(module, kind): oracle/_ref (the reference's unmodified sources) when its library is there or can be built,
    else the restated oracle.  Neither loads the product library."""
    from oracle import ref_py

    if ref_py.available():
        try:
            ref_py.build()
            return ref_py, "reference"
        except Exception:
            pass
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py, "port"


def cpu_solve(mod, kind, x0, trig, threads, instrument):
    if kind == "reference":
        return mod.ilqr_solve_batch(mod.MODEL_ST_LANE, x0, max_iterations=MAX_ITER, tolerance=TOL, trig=trig, threads=threads, instrument=instrument)
    return mod.ilqr_solve_batch(mod.MODEL_ST_LANE, x0, max_iterations=MAX_ITER, tolerance=TOL, trig=trig, threads=threads)


def cpu_describe(kind: str) -> str:
    if kind == "reference":
        return ("oracle/_ref/libref.so = the reference's unmodified headers + examples/single_track_ocp.cpp (g++ -O3 -DNDEBUG -fopenmp, "
                "std::function callbacks and heap temporaries included) against oracle/eigen_shim; glibc libm; one fresh iLQR per problem, "
                "`omp parallel for schedule(static)` over problems as in strategies/nash.hpp:59-64")
    return "oracle/ (C++ restatement of the reference, glibc libm, -O3 -ffp-contract=off, OpenMP static over problems); oracle/_ref not available"


def cpu_reference_run(n_problems: int, steps: int, warmup: int, threads: int = 0):
    """Times the reference's CPU implementation on the first n_problems problems of the workload.
    Returns (solves/s, threads, ms/step, kind)."""
    mod, kind = cpu_checker()
    x0 = mod.synthetic_single_track_x0(PROBLEMS)[:n_problems]
    # all host cores this process may use (torchrun sets OMP_NUM_THREADS=1; the num_threads clause overrides it)
    threads = threads or len(os.sched_getaffinity(0))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_solve(mod, kind, x0, 0, threads, instrument=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = float(np.mean(times))
    return n_problems / mean, threads, mean * 1e3, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mod, kind = cpu_checker()
    threads = len(os.sched_getaffinity(0))
    # size the per-step sample so that the whole --steps/--warmup run stays within ~2 minutes: probe the speed first
    x_probe = mod.synthetic_single_track_x0(PROBLEMS)[:2048]
    cpu_solve(mod, kind, x_probe[:256], 0, threads, instrument=False)
    t0 = time.perf_counter()
    cpu_solve(mod, kind, x_probe, 0, threads, instrument=False)
    rate = 2048 / max(time.perf_counter() - t0, 1e-6)
    warm = max(args.warmup, 1)
    n = args.cpu_sample if args.cpu_sample > 0 else int(min(PROBLEMS, max(4096, (110.0 * rate / (args.steps + warm)) // 4096 * 4096)))
    value, threads, ms, kind = cpu_reference_run(n, args.steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": shared_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"first {n} of the 65,536 problems per step (throughput per solve does not depend on the batch size); " + cpu_describe(kind)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- parity gate -------------------------------------------------------------------------------------------------
def parity_gate(x0, got, threads):
    """Results of the timed batch (rank 0's shard) against the CPU checkers.  Not timed."""
    mod, kind = cpu_checker()
    t0 = time.perf_counter()
    ref = cpu_solve(mod, kind, x0, 1, threads, instrument=True)  # portable trig: the mode the kernels are bit-compatible with
    n = x0.shape[0]
    same = (np.all(got["X"] == ref["X"], axis=(1, 2)) & np.all(got["U"] == ref["U"], axis=(1, 2)) & (got["cost"] == ref["cost"]))
    rel = np.abs(got["cost"] - ref["cost"]) / np.maximum(np.abs(ref["cost"]), 1e-300)
    out = {
        "checker": "oracle/_ref: the reference's unmodified sources, portable trig bound at link time" if kind == "reference"
                   else "oracle/ restatement (oracle/_ref not available), portable trig",
        "problems_checked": int(n), "bit_equal_problems": int(same.sum()),
        "max_rel_cost": float(rel.max()), "max_abs_dX": float(np.abs(got["X"] - ref["X"]).max()), "max_abs_dU": float(np.abs(got["U"] - ref["U"]).max()),
        "iterations_equal": bool(np.array_equal(got["iterations"], ref["iterations"])), "status_equal": bool(np.array_equal(got["status"], ref["status"])),
        "tolerances": {"cost_rel": 1e-9, "traj_abs": 1e-7},
    }
    out["pass"] = bool(out["max_rel_cost"] <= 1e-9 and out["max_abs_dX"] <= 1e-7 and out["max_abs_dU"] <= 1e-7 and out["iterations_equal"]
                       and out["status_equal"])
    # (ii) against the reference in its own libm mode, and the reference against itself under a 1-ulp change of x0: the
    # problem's finite-difference cross term (finite_differences.hpp:263-287) amplifies the last bit of libm
    g = cpu_solve(mod, kind, x0, 0, threads, instrument=True)
    relg = np.abs(got["cost"] - g["cost"]) / np.maximum(np.abs(g["cost"]), 1e-300)
    dxg = np.maximum(np.abs(got["X"] - g["X"]).max(axis=(1, 2)), np.abs(got["U"] - g["U"]).max(axis=(1, 2)))
    nb = min(n, 8192)
    x0p = x0[:nb].copy()
    x0p[:, 1] = np.nextafter(x0p[:, 1], np.inf)
    gp = cpu_solve(mod, kind, x0p, 0, threads, instrument=False)
    relb = np.abs(gp["cost"] - g["cost"][:nb]) / np.maximum(np.abs(g["cost"][:nb]), 1e-300)
    dxb = np.maximum(np.abs(gp["X"] - g["X"][:nb]).max(axis=(1, 2)), np.abs(gp["U"] - g["U"][:nb]).max(axis=(1, 2)))
    out["vs_reference_libm"] = {
        "frac_cost_within_1e-9": float((relg <= 1e-9).mean()), "frac_traj_within_1e-7": float((dxg <= 1e-7).mean()),
        "median_rel_cost": float(np.median(relg)), "max_rel_cost": float(relg.max()),
        "iterations_equal_frac": float((got["iterations"] == g["iterations"]).mean()), "status_equal_frac": float((got["status"] == g["status"]).mean()),
        "reference_own_1ulp_band": {"problems": int(nb), "frac_cost_within_1e-9": float((relb <= 1e-9).mean()),
                                    "frac_traj_within_1e-7": float((dxb <= 1e-7).mean()), "median_rel_cost": float(np.median(relb)),
                                    "max_rel_cost": float(relb.max()),
                                    "what": "the reference (glibc libm) against itself with x0[1] moved by one ulp"},
    }
    out["seconds"] = time.perf_counter() - t0
    return out


def bind_to_gpu_numa_node(torch, local_rank: int):
    """Pins this rank (threads and, by first touch, its pinned host buffers) to the NUMA node its GPU hangs off, when the host
    exposes one: results travel device -> host every step, and a rank whose buffers sit on the other socket pays the
    inter-socket link on all of them.  Returns what was done, for the JSON line."""
    info = {"numa_node": None, "bound": False}
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as exc:  # noqa: BLE001
        info["error"] = repr(exc)[:120]
    return info


class Lane:
    """One solve pipeline: its own CUDA stream, engine context, resident batch and pinned result buffers."""

    def __init__(self, torch, mas, device, desc, per_rank, args, with_host_buffers, priority=0):
        self.stream = torch.cuda.Stream(priority=priority)
        self.ctx = mas.Context(device, self.stream.cuda_stream)
        if args.blocking_sync:
            self.ctx.set_blocking_sync(True)
        self.batch = mas.Batch(self.ctx, desc, per_rank)
        if args.lanes or args.chains:
            self.batch.set_tuning(args.lanes, args.chains)
        if args.ls_mode:
            self.batch.set_line_search_mode(args.ls_mode)
        if args.hint:
            self.batch.set_concurrency_hint(args.hint)
        self.out = None
        if with_host_buffers:
            self.out = dict(X=torch.empty((per_rank, T + 1, NX), dtype=torch.float64).pin_memory().numpy(),
                            U=torch.empty((per_rank, T, NU), dtype=torch.float64).pin_memory().numpy(),
                            cost=torch.empty(per_rank, dtype=torch.float64).pin_memory().numpy(),
                            iterations=torch.empty(per_rank, dtype=torch.int32).pin_memory().numpy(),
                            status=torch.empty(per_rank, dtype=torch.int32).pin_memory().numpy())

    def close(self):
        self.batch.close()


def best_of(fn, repeats=3):
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def circle_scenarios(n_scen, n_agents, seed=0, jitter=True):
    rng = np.random.default_rng(seed)
    th = 2.0 * np.pi * np.arange(n_agents) / n_agents
    R = rng.uniform(15, 25, (n_scen, 1)) if jitter else np.full((n_scen, 1), 20.0)
    x0 = np.stack([R * np.cos(th), R * np.sin(th), np.broadcast_to(1.57 + th, (n_scen, n_agents)), np.full((n_scen, n_agents), 4.0)], -1)
    gp = np.broadcast_to(np.stack([R[:, 0], np.full(n_scen, 5.0), np.ones(n_scen), np.ones(n_scen), np.full(n_scen, 1e-3), np.full(n_scen, 1e-3)],
                                  -1)[:, None, :], (n_scen, n_agents, 6)).copy()
    return x0, gp


def run_other_configs(mas, torch, dist, rank, world, local_rank):
    """Short runs of BASELINE configs 1, 2, 4, 5 through the C ABI with host buffers (wall clock around synchronous
    calls, best of a few).  Every block is independent; a failure is recorded, never fatal for the headline line."""
    out = {}
    ctx = mas.Context(local_rank)
    p10, p100 = mas.IlqrParams.make(10, 1e-5), mas.IlqrParams.make(100, 1e-5)

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:  # noqa: BLE001
            out[name] = {"error": repr(exc)[:300]}

    def config1():
        b = mas.Batch(ctx, mas.example_desc(0), 1)
        x1 = np.array([[0.0, 1.0, 0.0, 0.0]])

        def solve1():
            b.set_initial_states(x1)
            b.set_controls(None)
            b.solve(p10)
            return b.get_solution()

        t = best_of(solve1, 7)
        r = solve1()
        b.close()
        return {"what": "single_track_ocp --solver ilqr, one problem, single-solve latency (host buffers)", "latency_us": t * 1e6,
                "cost": float(r["cost"][0]), "iterations": int(r["iterations"][0])}

    def config2():
        S, A = 4096, 3
        x0, gp = circle_scenarios(S, A)
        d1 = mas.example_desc(1)
        t = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, d1, p100, 10, x0, model_params=gp, trace=False), 3)
        return {"what": "multi_agent_single_track --agents 3 --strategy trustregion x 4,096 scenarios (radius jittered), 10 outer rounds, one GPU",
                "scenarios_per_s": S / t, "ms": t * 1e3}

    def config5():
        d1 = mas.example_desc(1)
        th = 2.0 * np.pi * np.arange(32) / 32
        xe = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
        rec = {"what": "multi_agent_single_track --agents 32 --strategy centralized (stacked n=128, m=64, all-FD), one GPU"}
        t1 = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, xe, trace=False), 3)
        rec["one_scenario_ms"] = t1 * 1e3
        S5 = 592
        x5 = np.repeat(xe, S5, axis=0)
        t = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x5, trace=False), 2)
        rec["replicas"] = S5
        rec["scenarios_per_s"] = S5 / t
        # track radius jittered per scenario: iteration counts spread from 4 to ~90, one scenario is one CTA, so a run is bounded
        # below by its longest scenario; persistent CTAs pull scenarios from a queue, and with enough scenarios the rate
        # approaches (SMs / mean scenario time)
        for Sj in (296, 1184):
            xj, gpj = circle_scenarios(Sj, 32, seed=5)
            rj = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, xj, model_params=gpj)
            tj = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, xj, model_params=gpj, trace=False), 1)
            its = rj["trace_iters"][:, 0, 0]
            rec[f"jittered_radius_{Sj}"] = {"scenarios_per_s": Sj / tj, "ms": tj * 1e3, "iterations_mean": float(its.mean()), "iterations_max": int(its.max()),
                                            "longest_scenario_bound_ms": float(its.max()) * t1 * 1e3 / 4.0,
                                            "work_bound_ms": float(its.sum()) * t1 * 1e3 / 4.0 / 148.0}
        return rec

    if rank == 0:
        guarded("config1", config1)
        guarded("config2", config2)
        guarded("config5", config5)

    # config 4: 1,024 LQR agents, sequential, 10 outer rounds; at N > 1 the agents are sharded over the ranks (1,024 / N
    # each) and every round ends with the library's NCCL all-gather of all agents' (X, U, cost)
    def config4():
        A_total, outer = 1024, 10
        A_loc = A_total // world
        x4 = np.tile([1.0, 0.0, 0.0, 0.0], (1, A_total, 1))  # multi_agent_lqr.cpp:30-33: every agent starts at (1, 0, 0, 0)
        d2 = mas.example_desc(2)
        cctx = mas.Context(local_rank)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.frombuffer(bytearray(mas.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(uid, 0)
            cctx.init_nccl(bytes(uid.cpu().numpy().tobytes()), rank, world)
            cctx.set_agent_sharding(True)
        mine = x4[:, rank * A_loc:(rank + 1) * A_loc]
        res = {}

        def run():
            res["r"] = mas.strategy_run(cctx, mas.Strategy.SEQUENTIAL, d2, p100, outer, mine, trace=False)

        run()
        ts, cs = [], []
        for _ in range(3):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
            cs.append(cctx.exchange_stats()["collective_ms"] if world > 1 else 0.0)
        t = min(ts)
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rec = {"what": f"multi_agent_lqr --agents 1024 --strategy sequential, 10 outer rounds, {A_loc} agents per GPU",
               "ms": float(tt[0]) * 1e3, "agent_solves_per_s": A_total * outer / float(tt[0]), "total_cost": float(res["r"]["total_cost"][0])}
        if world > 1:
            st = cctx.exchange_stats()
            rec["allgather"] = {"per_round_ms": min(cs) / max(st["rounds"], 1), "rounds": st["rounds"], "bytes_received_per_round": st["bytes_per_round"],
                                "what": "ncclAllGather of every agent's (X, U, cost) after every outer round, CUDA events around the group"}
            # the same agents on one GPU without a communicator: the joint total must be the same bits
            if rank == 0:
                full = mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, d2, p100, outer, x4, trace=False)
                rec["bit_equal_to_one_gpu"] = bool(full["total_cost"][0] == res["r"]["total_cost"][0]
                                                   and np.array_equal(full["costs"][0, :A_loc], res["r"]["costs"][0]))
        cctx.close()
        return rec

    try:
        out["config4_nash"] = config4()
    except Exception as exc:  # noqa: BLE001
        out["config4_nash"] = {"error": repr(exc)[:300]}
    ctx.close()
    return out


def run_b200(args):
    import torch
    import multi_agent_solver_b200 as mas

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if (world > 1 and not args.no_numa_bind) else {"numa_node": None, "bound": False}
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    scaling = args.scaling if args.scaling != "auto" else ("strong" if world > 1 else "weak")
    x0_all = mas.synthetic_single_track_x0(PROBLEMS)
    desc = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
    prm = mas.IlqrParams.make(MAX_ITER, TOL)
    # solves in flight per GPU: a full 65,536-problem shard fills the device in its first iterations (4 pipelines hide the
    # latency-bound tail); the smaller shards of strong scaling are latency-bound from the start, so more of them are kept
    # in flight, and every batch is told that it shares the device (mas_b200_batch_set_concurrency_hint: narrower lane
    # mappings, i.e. less speculative work).  Measured on one B200 (profiles/r02_strong_scaling_tuning.txt).
    shard = args.shard or (PROBLEMS if scaling == "weak" else PROBLEMS // world)

    def depth_for(per_rank: int) -> int:
        """Solves in flight per GPU.  A full 65,536-problem shard fills the device in its first iterations and four pipelines
        hide its latency-bound tail.  The smaller shards of strong scaling are latency-bound from the start (every
        iteration costs the latency of its 80 sequential time steps whatever the problem count), so the device only fills
        with many of them in flight: up to 12 for 32,768 problems, up to 24 below -- never more than there are steps.
        Measured on one B200 (profiles/r02_strong_scaling_tuning.txt)."""
        if args.depth > 0:
            return args.depth
        if args.steps < 8:
            return min(3, max(args.steps, 1))
        cap = 4 if per_rank >= PROBLEMS else (12 if per_rank >= PROBLEMS // 2 else 24)
        return max(1, min(cap, args.steps))

    depth = depth_for(shard)
    if args.hint == 0 and shard < PROBLEMS:
        args.hint = 4  # every batch is told that it shares the device: narrower lane mappings, less speculative work
    # more host threads than cores (8 ranks x many pipelines on a 32-vCPU box): wait on events sleeping, not spinning
    cores = len(os.sched_getaffinity(0))
    if args.blocking_sync < 0:
        args.blocking_sync = 1 if world * (depth + 1) > cores else 0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(mode: str, steps: int, full: bool):
        """One scaling mode: strong = the 65,536 problems split over the ranks, weak = 65,536 per rank."""
        per_rank = PROBLEMS if mode == "weak" else PROBLEMS // world
        if args.shard:  # profiling aid: the shard a rank would hold at N = 65,536 / shard GPUs, on this GPU alone
            per_rank = args.shard
        total = per_rank * world
        if mode == "weak":
            # every rank gets the 65,536 problems, rotated so the ranks do not solve identical shards in the same order
            x0 = np.roll(x0_all, -rank * 4099, axis=0).copy()
        else:
            x0 = x0_all[(rank * per_rank) % PROBLEMS:(rank * per_rank) % PROBLEMS + per_rank].copy()
        x0_host = torch.from_numpy(x0).pin_memory().numpy()
        n_lanes = depth_for(per_rank)
        # e2e results: downloaded after the solve by the copy engine (default), or streamed into a result sink while the solve
        # runs (--sink 1; mas_b200_batch_set_result_sink).  Measured both ways (profiles/r02_strong_scaling_tuning.txt): the
        # sink wins on a single-GPU host when pipelines run one or two steps each, the copy engine wins for 65,536-problem
        # shards and on the multi-GPU hosts, where SM stores of several GPUs reach less host bandwidth than their copy engines
        use_sink = args.sink > 0

        def make_lanes(n, host_buffers, prio):
            # prio: stream priorities falling with the lane index, so that deep pipelines complete their solves one after the
            # other instead of all at the end -- a finished lane's download then overlaps the other lanes' solves instead of
            # queueing on the copy engine behind everyone else's after the last kernel (e2e of small shards +17-21 %; the
            # resident measurement loses 12 % with it and does not use it)
            lo, hi = (0, -5) if prio else (0, 0)  # B200: cudaDeviceGetStreamPriorityRange = [0, -5]; torch clamps
            return [Lane(torch, mas, local_rank, desc, per_rank, args, with_host_buffers=host_buffers, priority=hi + (i * (lo - hi + 1)) // n)
                    for i in range(n)]

        lanes = make_lanes(n_lanes, False, args.priorities > 0)
        for ln in lanes:
            ln.batch.set_initial_states(x0_host)  # resident input of the `value` measurement

        def resident_step(ln):
            ln.batch.set_controls(None)  # U_init = 0 (device memset); x0 is already resident
            ln.batch.solve(prm)

        def e2e_step(ln, keys=None):
            ln.batch.set_initial_states(x0_host)  # H2D from pinned memory
            ln.batch.set_controls(None)
            ln.batch.solve(prm)
            # D2H of X, U, cost, iterations, status into pinned memory.  Default: the lane's pinned buffers are registered as the
            # batch's result sink (mas_b200_batch_set_result_sink) and the solve itself streams every problem's rows out as soon as
            # the problem leaves the active set -- the transfer overlaps the remaining iterations.  --no-sink: staged in HBM on the
            # solve stream after the solve and copied on the batch's copy stream (mas_b200_batch_begin_get_solution).  Either way
            # e2e_finish() waits for the lane's last results inside the timed region.
            if use_sink:
                return
            ln.batch.begin_get_solution(ln.out if keys is None else {k: v for k, v in ln.out.items() if k in keys})

        def e2e_finish(ln):
            ln.batch.wait_solution()

        def set_sinks(use_lanes, keys=None):
            if not use_sink:
                return
            for ln in use_lanes:
                ln.batch.set_result_sink(ln.out if keys is None else {k: v for k, v in ln.out.items() if k in keys})

        def timed(fn, nsteps, use_lanes, stagger_ms, finish=None, stagger_units=None):
            """Runs exactly `nsteps` steps spread over `use_lanes` pipelines, each driven by its own host thread and
            started stagger_ms/len(use_lanes) apart so one solve's latency-bound tail overlaps another's bulk.
            Device time = latest end event - earliest start event over the lanes' streams (then max over ranks)."""
            n = len(use_lanes)
            counts = [nsteps // n + (1 if i < nsteps % n else 0) for i in range(n)]
            e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
            go = threading.Barrier(n + 1)
            errors = []

            def work(i):
                try:
                    torch.cuda.set_device(local_rank)  # the current device is per host thread
                    ln = use_lanes[i]
                    go.wait()
                    # starts staggered when every pipeline runs several steps (one solve's latency-bound tail then overlaps
                    # another's bulk for the whole run); with one or two steps per pipeline the offsets would only delay the
                    # last start, so they all start together (measured: profiles/r02_strong_scaling_tuning.txt)
                    stagger = stagger_units if stagger_units is not None else (args.stagger if args.stagger >= 0 else (1.0 if nsteps > 2 * n else 0.0))
                    if i and stagger > 0:
                        time.sleep(i * stagger_ms * stagger * 1e-3 / n)
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
            return ms / nsteps, wall / nsteps

        res = {"per_rank": per_rank, "total": total, "x0": x0, "lanes": lanes, "depth": n_lanes, "sink": use_sink}
        # ---- one solve at a time: latency of one solve of the shard, and the reference for kernel shares -----------
        for _ in range(max(args.warmup, 3)):
            for ln in lanes:
                resident_step(ln)
        barrier()
        res["single_ms"], _ = timed(resident_step, max(3, min(steps, 6)), lanes[:1], 0.0)
        # ---- resident throughput: `depth` solves in flight ------------------------------------------------------
        launches0 = sum(ln.batch.stats()["kernel_launches"] for ln in lanes)
        res["ms_step"], res["wall_step"] = timed(resident_step, steps, lanes, res["single_ms"])
        res["launches"] = sum(ln.batch.stats()["kernel_launches"] for ln in lanes) - launches0
        res["value"] = total / (res["ms_step"] * 1e-3)
        res["stats"] = lanes[0].batch.stats()
        if not full or args.resident_only:
            return res
        # ---- e2e: host buffers in, host buffers out, every step ---------------------------------------------------
        # its own pipelines: pinned result buffers, and for the small shards of strong scaling fewer of them than the resident
        # measurement keeps in flight.  At N > 1 the e2e run is bound by the host's aggregate device-to-host bandwidth (255 MB per
        # step whatever N: profiles/r02_pcie_probe_*), so what counts is that the first results start travelling early and the copy
        # engines never idle -- 6 / 8 staggered pipelines did better than 12 / 20 with or without stream priorities (2 and 8 GPUs)
        n_e2e = args.e2e_depth if args.e2e_depth > 0 else (n_lanes if per_rank >= PROBLEMS else min(n_lanes, 6 if per_rank >= PROBLEMS // 2 else 8))
        e2e_prio = args.priorities > 0
        for ln in lanes[1:]:
            ln.close()
        e2e_lanes = make_lanes(n_e2e, True, e2e_prio)
        lanes = lanes[:1]
        res["lanes"] = lanes + e2e_lanes
        for ln in e2e_lanes:
            ln.batch.set_initial_states(x0_host)
        set_sinks(e2e_lanes)
        for ln in e2e_lanes:
            e2e_step(ln)
            e2e_finish(ln)
        # end to end the pipelines always start staggered: a lane's download then overlaps the other lanes' solves
        e2e_stagger = args.e2e_stagger if args.e2e_stagger >= 0 else 1.0
        res["e2e_depth"], res["e2e_priorities"] = n_e2e, bool(e2e_prio)
        e2e_ms, e2e_wall = timed(e2e_step, steps, e2e_lanes, res["single_ms"], finish=e2e_finish, stagger_units=e2e_stagger)
        res["e2e"] = {"value": total / (max(e2e_ms * 1e-3, e2e_wall)), "unit": UNIT, "h2d_bytes_per_step": per_rank * NX * 8,
                      "d2h_bytes_per_step": per_rank * (((T + 1) * NX + T * NU + 1) * 8 + 2 * 4), "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall * 1e3,
                      "results_via": "result sink: rows stored to pinned host memory by the engine as problems finish (mas_b200_batch_set_result_sink)"
                                     if use_sink else "staged in HBM after the solve, copy engine on a second stream (mas_b200_batch_begin_get_solution)"}
        # what the timed batch produced (lane 0's last download): input of the parity gate
        res["timed_output"] = {k: v.copy() for k, v in e2e_lanes[0].out.items()}
        # the same loop for a caller that only wants the controls (X = NULL): a third of the download, reported next to the
        # headline e2e because at N > 1 the host's D2H bandwidth sets e2e
        ukeys = ("U", "cost", "iterations", "status")
        set_sinks(e2e_lanes, ukeys)
        for ln in e2e_lanes:
            e2e_step(ln, ukeys)
            e2e_finish(ln)
        u_ms, u_wall = timed(lambda ln: e2e_step(ln, ukeys), steps, e2e_lanes, res["single_ms"], finish=e2e_finish, stagger_units=e2e_stagger)
        res["e2e_controls_only"] = {"value": total / (max(u_ms * 1e-3, u_wall)), "unit": UNIT, "h2d_bytes_per_step": per_rank * NX * 8,
                                    "d2h_bytes_per_step": per_rank * ((T * NU + 1) * 8 + 2 * 4), "ms_per_step": u_ms,
                                    "note": "same loop without downloading the state trajectories (X = NULL in the C ABI); not the headline"}
        set_sinks(e2e_lanes, ())  # unregister
        # ---- per-kernel timing for the roofline: one solve at a time, CUDA events around every launch inside the engine
        batch = lanes[0].batch
        batch.set_profiling(True)
        for _ in range(max(2, min(steps, 5))):
            resident_step(lanes[0])
        res["prof"] = batch.profile()
        batch.set_profiling(False)
        res["fp64_peak"] = lanes[0].ctx.probe_fp64_peak()
        return res

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main = measure(scaling, args.steps, full=True)
    clocks = sampler.stop() if rank == 0 else None
    for ln in main["lanes"]:
        ln.close()

    if args.resident_only:  # short form for ncu captures: only the resident steps above
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": main["value"], "unit": UNIT, "ms_per_step": main["ms_step"], "single_solve_ms": main["single_ms"],
                              "gpu_launches": int(main["launches"]), "note": "resident-only run (profiling aid), not a bench line"}), flush=True)
        return

    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        w = measure("weak", args.steps, full=False)
        weak = {"value": w["value"], "unit": UNIT, "ms_per_step": w["ms_step"], "problems_per_gpu": w["per_rank"], "problems_total": w["total"],
                "what": "one full 65,536-problem shard per rank, resident inputs (round 1's N > 1 headline)"}
        for ln in w["lanes"]:
            ln.close()

    other = None
    if not args.no_configs:
        # the library's own NCCL communicator may announce its version on stdout: keep stdout clean for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            other = run_other_configs(mas, torch, dist, rank, world, local_rank)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    if rank == 0:
        per_rank, total = main["per_rank"], main["total"]
        st, prof, single_ms, ms_step = main["stats"], main["prof"], main["single_ms"], main["ms_step"]
        out = main["timed_output"]
        threads = len(os.sched_getaffinity(0))
        parity = None
        if not args.no_parity:
            try:
                parity = parity_gate(main["x0"], out, threads)
            except Exception as exc:  # noqa: BLE001
                parity = {"error": repr(exc)[:300]}
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
        hbm_achieved = (k["alg_bytes_per_solve"] / n_launch) / (k["ms_per_solve"] / n_launch * 1e-3) / 1e9
        # algorithmic flops of one solve of the shard (SURVEY 8d convention), split by kernel: trial rollouts + the accepted
        # step's rollout for the line search, derivatives + Riccati for the backward pass
        fwd_flops = T * (st["alpha_trials"] + st["iterations"]) * FWD_FLOPS_STEP
        bwd_flops = T * st["iterations"] * BWD_FLOPS_STEP
        alg_flops = fwd_flops + bwd_flops + T * per_rank * FWD_FLOPS_STEP
        dom_flops = fwd_flops if dom == "forward_kernel" else bwd_flops
        fp64_peak = main["fp64_peak"]
        dom_tflops = dom_flops / (k["ms_per_solve"] * 1e-3) / 1e12
        traffic, traffic_src = load_traffic(dom)
        roofline = {
            "bound": "fp64", "kernel": "line search (forward_coop_kernel + forward_kernel)" if dom == "forward_kernel" else dom,
            "achieved": dom_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": dom_tflops / fp64_peak if fp64_peak else None,
            "peak_source": "DFMA throughput probe on this GPU (mas_b200_probe_fp64_peak: 8 independent FMA chains per thread, 8 x 256 threads "
                           "per SM, CUDA events); MEASURED_PEAKS.json holds no fp64 figure.  The kernels run -fmad=false (the reference build has no "
                           "FMA contraction), so a multiply-add is two pipe slots: the ceiling of `frac` for unfused arithmetic is 0.5",
            "achieved_what": "algorithmic flops of the sequential reference for this kernel's share of one solve (SURVEY 8d convention: "
                             "MAC = 2, one trig / division = 32) / the kernel's summed launch time, one solve in flight",
            "traffic": traffic, "traffic_source": traffic_src,
            "avg_launch_ms": k["ms_per_solve"] / n_launch, "alg_bytes_per_launch": k["alg_bytes_per_solve"] / n_launch,
            "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                    "note": "algorithmic bytes / launch time: HBM is not the bound of this kernel"},
            "kernel_share_of_single_solve": {name: v["ms_per_solve"] / single_ms for name, v in kernels.items()},
            "whole_step_fp64": {"alg_tflops": alg_flops / (ms_step * 1e-3) / 1e12,
                                "frac": alg_flops / (ms_step * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                                "note": f"all kernels of a step / ms_per_step with {depth} solves in flight"},
        }
        cpu_baseline = None
        if n_gpus == 1:  # reported on rank 0 at N=1 only
            sample = args.cpu_sample if args.cpu_sample > 0 else 16384
            cpu_val, cpu_threads, cpu_ms, kind = cpu_reference_run(sample, 1, 1)
            cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": cpu_threads, "kind": kind,
                            "sample": f"first {sample} of the 65,536 problems, one pass ({cpu_ms:.0f} ms) after one warm-up pass; " + cpu_describe(kind)}
        cfg = shared_config(n_gpus)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "engine": {"problems_per_gpu": per_rank, "parallelism": f"contiguous shards of independent problems x{n_gpus}, no data-path collective",
                       "solves_in_flight": depth, "e2e_solves_in_flight": main.get("e2e_depth"), "e2e_stream_priorities": main.get("e2e_priorities"),
                       "cuda_device_max_connections": int(os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "8")),
                       "concurrency_hint": args.hint or 1, "blocking_sync": bool(args.blocking_sync),
                       "pipelining": f"{depth} independent solves of the shard in flight per GPU, each a whole step on its own stream and host "
                                     "thread (starts staggered when a pipeline runs more than two steps); ms_per_step = device time of the K steps / K",
                       "single_solve_ms": single_ms, "host_cores": len(os.sched_getaffinity(0)), "numa": numa,
                       "l2": f"working set {per_rank * 10272 / 1e6:.0f} MB per solve (X,U,K,k) x {depth} in flight"
                             + (" > 126 MB L2, no flush needed" if per_rank * 10272 * depth > 126e6 else " (fits L2: flushed by the other lanes' traffic only)"),
                       "forward_lanes": st["forward_lanes"], "forward_chains": st["forward_chains"],
                       "mean_iterations": iters_total / per_rank, "problem_iterations_per_s": main["value"] * iters_total / per_rank},
            "e2e": main["e2e"], "e2e_controls_only": main["e2e_controls_only"],
            "gpu_launches": int(main["launches"]),
            "parity": parity,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "weak": weak,
            "configs": other,
            "clocks": clocks,
            "wall_ms_per_step": main["wall_step"] * 1e3,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48, help="timed steps (one step = one solve of the 65,536-problem batch, 6-7 ms)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"], help="auto = strong (65,536 split over the ranks) at N > 1")
    ap.add_argument("--depth", type=int, default=0, help="independent solves in flight per GPU (1 = one at a time; 0 = by shard size and step count: "
                    "4 for a 65,536-problem shard, up to 12 / 24 for the 32,768 / smaller shards of strong scaling, never more than --steps)")
    ap.add_argument("--stagger", type=float, default=-1.0, help="start offset between pipelines, in units of single_solve_ms / depth (-1 = auto: 1 when "
                    "every pipeline runs more than two steps, else 0)")
    ap.add_argument("--sink", type=int, default=-1, help="e2e results: 1 = streamed into a result sink while the solve runs (mas_b200_batch_set_result_sink), "
                    "0 / -1 = downloaded after the solve by the copy engine (mas_b200_batch_begin_get_solution; default, measured faster on the multi-GPU hosts)")
    ap.add_argument("--priorities", type=int, default=-1, help="1: stream priorities falling with the pipeline index (an experiment: +17-21 % e2e for deep pipelines of small shards on a "
                    "single-GPU host, -12 % resident; profiles/r02_strong_scaling_tuning.txt); default off")
    ap.add_argument("--e2e-depth", type=int, default=0, help="pipelines used by the e2e measurement (0 = all)")
    ap.add_argument("--e2e-stagger", type=float, default=-1.0, help="start offset between pipelines of the e2e measurement (-1 = 1.0)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems per CPU-baseline pass (0 = sized automatically)")
    ap.add_argument("--resident-only", action="store_true", help="run only warm-up + timed resident steps (for ncu)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 1/2/4/5 blocks")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling measurement at N > 1")
    ap.add_argument("--ls-mode", type=int, default=0, help="line search scheduling: 0 auto, 1 concurrent lanes, 2 compacted rounds")
    ap.add_argument("--shard", type=int, default=0, help="profiling aid (with --resident-only): problems per rank, overriding 65,536 / N")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin ranks to the NUMA node of their GPU (N > 1)")
    ap.add_argument("--blocking-sync", type=int, default=-1, help="1: pipelines wait on events sleeping (mas_b200_context_set_blocking_sync); -1 = when oversubscribed")
    ap.add_argument("--hint", type=int, default=0, help="mas_b200_batch_set_concurrency_hint of every pipeline (0 = leave at 1)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--chains", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
