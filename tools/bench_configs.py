"""Secondary measurements: BASELINE.json configs 1, 2, 4, 5 (the headline config 3 is bench.py).

Each config is run through the C ABI with host buffers (what a caller of the reference's API would see), timed
with the wall clock around synchronous calls, best of several repeats after a warm-up (the shared boxes show sporadic
10x outliers on these short host-driven loops), next to the oracle (CPU
restatement of the reference) on a bounded sample with all host threads.  One JSON line per config.

    python tools/bench_configs.py > gpurun_out/other_configs.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402
from oracle import oracle_py as o  # noqa: E402


def best_of(fn, repeats=3):
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def circle(n_scen, n_agents, seed=0):
    rng = np.random.default_rng(seed)
    th = 2.0 * np.pi * np.arange(n_agents) / n_agents
    R = rng.uniform(15, 25, (n_scen, 1))
    x0 = np.stack([R * np.cos(th), R * np.sin(th), np.broadcast_to(1.57 + th, (n_scen, n_agents)), np.full((n_scen, n_agents), 4.0)], -1)
    gp = np.broadcast_to(np.stack([R[:, 0], np.full(n_scen, 5.0), np.ones(n_scen), np.ones(n_scen), np.full(n_scen, 1e-3), np.full(n_scen, 1e-3)],
                                  -1)[:, None, :], (n_scen, n_agents, 6)).copy()
    op = gp[:, :, :2].copy()
    return x0, gp, op


def main():
    ctx = mas.Context(0)
    threads = len(os.sched_getaffinity(0))
    out = []

    # config 1: the single_track_ocp problem, single-solve latency
    d0 = mas.example_desc(0)
    p10 = mas.IlqrParams.make(10, 1e-5)
    x1 = np.array([[0.0, 1.0, 0.0, 0.0]])
    b = mas.Batch(ctx, d0, 1)

    def solve1():
        b.set_initial_states(x1)
        b.set_controls(None)
        b.solve(p10)
        return b.get_solution()

    t_gpu = best_of(solve1, 5)
    r = solve1()
    t_cpu = best_of(lambda: o.ilqr_solve_batch(0, x1, max_iterations=10, tolerance=1e-5, threads=1), 5)
    out.append({"config": 1, "what": "single_track_ocp --solver ilqr, one problem, latency", "gpu_us": t_gpu * 1e6, "cpu_oracle_us": t_cpu * 1e6,
                "cost": float(r["cost"][0]), "iterations": int(r["iterations"][0])})
    b.close()

    # config 2: 4,096 scenarios x 3 agents, trust region, 10 outer rounds
    S, A = 4096, 3
    x0, gp, op = circle(S, A)
    d1 = mas.example_desc(1)
    p100 = mas.IlqrParams.make(100, 1e-5)
    t_gpu = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, d1, p100, 10, x0, model_params=gp, trace=False), 6)
    ns = 256
    t_cpu = best_of(lambda: o.strategy_run_batch(o.STRATEGY_TRUSTREGION, 1, x0[:ns], params=op[:ns], max_outer=10, max_iterations=100, tolerance=1e-5,
                                                 threads=threads), 1)
    out.append({"config": 2, "what": "multi_agent_single_track --agents 3 --strategy trustregion x 4,096 scenarios (radius jittered), 10 outer rounds",
                "gpu_scenarios_per_s": S / t_gpu, "gpu_ms": t_gpu * 1e3, "cpu_oracle_scenarios_per_s": ns / t_cpu, "cpu_threads": threads,
                "cpu_sample_scenarios": ns})

    # config 4: 1,024 LQR agents, sequential, 10 outer rounds
    x4 = np.tile([1.0, 0.0, 0.0, 0.0], (1, 1024, 1))
    d2 = mas.example_desc(2)
    t_gpu = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, d2, p100, 10, x4, trace=False), 6)
    t_cpu = best_of(lambda: o.strategy_run_batch(o.STRATEGY_SEQUENTIAL, 2, x4, max_outer=10, max_iterations=100, tolerance=1e-5, threads=threads), 1)
    out.append({"config": 4, "what": "multi_agent_lqr --agents 1024 --strategy sequential, 10 outer rounds (one scenario)", "gpu_ms": t_gpu * 1e3,
                "cpu_oracle_ms": t_cpu * 1e3, "cpu_threads": threads, "gpu_agent_solves_per_s": 1024 * 10 / t_gpu})

    # config 5: 32 stacked agents (n = 128, m = 64), centralized, batch of scenarios
    # (a) the example's scenario (R = 20) replicated: every CTA does the same 4 iterations
    th = 2.0 * np.pi * np.arange(32) / 32
    xe = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
    for S5 in (1, 148, 592):
        x5 = np.repeat(xe, S5, axis=0)
        t_gpu = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x5, trace=False), 5)
        rec = {"config": 5, "what": f"multi_agent_single_track --agents 32 --strategy centralized, the example scenario x {S5} replicas "
                                    "(stacked n=128, m=64, all-FD, 4 iterations each)", "gpu_ms": t_gpu * 1e3, "gpu_scenarios_per_s": S5 / t_gpu}
        if S5 == 1:
            t0 = time.perf_counter()
            o.strategy_run_batch(o.STRATEGY_CENTRALIZED, 1, x5, max_outer=1, max_iterations=100, tolerance=1e-5, threads=1)
            rec["cpu_oracle_ms_one_scenario"] = (time.perf_counter() - t0) * 1e3
        out.append(rec)
    # (b) track radius jittered per scenario: iteration counts spread widely and the slowest scenario sets the time
    x5, gp5, _ = circle(296, 32, seed=5)
    r5 = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x5, model_params=gp5)
    t_gpu = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x5, model_params=gp5, trace=False), 2)
    its = r5["trace_iters"][:, 0, 0]
    out.append({"config": 5, "what": "same, 296 scenarios with the track radius jittered in [15, 25]", "gpu_ms": t_gpu * 1e3,
                "gpu_scenarios_per_s": 296 / t_gpu, "iterations_mean": float(its.mean()), "iterations_max": int(its.max())})

    # the pendulum swing-up (all-FD, time-varying cost, up to 1000 iterations) and the rocket
    # (analytic, mass clamp, 25 iterations) at batch scale, initial states jittered around the examples' own
    rng = np.random.default_rng(11)
    for model, name, B, x0 in (
            (3, "pendulum_swing_up", 4096, np.stack([np.pi - 0.05 + rng.uniform(-0.1, 0.1, 4096), rng.uniform(-0.1, 0.1, 4096)], -1)),
            (4, "rocket_max_altitude", 16384, np.stack([rng.uniform(0, 1, 16384), rng.uniform(-1, 1, 16384), 50.0 + rng.uniform(-2, 2, 16384)], -1))):
        max_it, tol = {3: (1000, 1e-4), 4: (25, 1e-6)}[model]
        dm = mas.example_desc(model)
        pm_ = mas.IlqrParams.make(max_it, tol)
        bb = mas.Batch(ctx, dm, B)

        def solve_b():
            bb.set_initial_states(x0)
            bb.set_controls(mas.example_controls(model, dm.horizon_steps)[None].repeat(B, 0))
            bb.solve(pm_)
            return bb.get_solution()

        t_gpu = best_of(solve_b, 2)
        rb = solve_b()
        ns = 128 if model == 3 else 2048
        t_cpu = best_of(lambda: o.ilqr_solve_batch(model, x0[:ns], max_iterations=max_it, tolerance=tol, threads=threads), 1)
        out.append({"config": name, "what": f"{name} x {B} problems, iLQR {max_it} / {tol:g}", "gpu_ms": t_gpu * 1e3, "gpu_solves_per_s": B / t_gpu,
                    "cpu_oracle_solves_per_s": ns / t_cpu, "cpu_threads": threads, "cpu_sample": ns,
                    "iterations_mean": float(rb["iterations"].mean()), "iterations_max": int(rb["iterations"].max())})
        bb.close()

    for rec in out:
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
