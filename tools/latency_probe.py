"""Small-batch latency and FD-heavy throughput, backward pass mappings side by side (A/B through
mas_b200_batch_set_backward_mode / MAS_B200_BACKWARD_MODE).  One JSON line per case.

    python tools/latency_probe.py > gpurun_out/latency_probe.jsonl
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402


def best_of(fn, repeats=5):
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ctx = mas.Context(0)
    x_all = mas.synthetic_single_track_x0(65536)
    p10 = mas.IlqrParams.make(10, 1e-5)
    d0 = mas.example_desc(0)
    for B in (1, 1024, 8192, 65536):
        x0 = np.array([[0.0, 1.0, 0.0, 0.0]]) if B == 1 else x_all[:B]
        for mode, tpmax in ((1, 0), (3, 0), (0, 0)):
            b = mas.Batch(ctx, d0, B)
            b.set_backward_mode(mode, tpmax)
            b.set_initial_states(x0)

            def solve():
                b.set_controls(None)
                b.solve(p10)
                ctx.synchronize()

            t = best_of(solve, 7)
            print(json.dumps({"case": f"ST-lane x {B}, resident, one solve", "backward_mode": mode, "ms": t * 1e3}), flush=True)
            b.close()
    # all-FD circular-track agents (config 2's path): 12,288 agents, one solve from zero controls
    rng = np.random.default_rng(0)
    th = rng.uniform(0, 2 * np.pi, 12288)
    xc = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(12288, 4.0)], -1)
    d1 = mas.example_desc(1)
    p100 = mas.IlqrParams.make(100, 1e-5)
    for mode in (1, 2, 3, 0):
        b = mas.Batch(ctx, d1, 12288)
        b.set_backward_mode(mode)
        b.set_initial_states(xc)

        def solve():
            b.set_controls(None)
            b.solve(p100)
            ctx.synchronize()

        t = best_of(solve, 5)
        print(json.dumps({"case": "ST-circ all-FD x 12,288, resident, one solve", "backward_mode": mode, "ms": t * 1e3}), flush=True)
        b.close()


if __name__ == "__main__":
    main()
