"""Experiment: several 65,536-problem solves in flight on one GPU, each on its own stream and driven by its own
host thread, staggered so that the latency-bound tail of one solve overlaps the bulk of another.
    python tools/pipeline_experiment.py [depths...]
"""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402


def main():
    depths = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    B = 65536
    x0 = mas.synthetic_single_track_x0(B)
    desc = mas.example_desc(0)
    prm = mas.IlqrParams.make(10, 1e-5)
    steps_total = 24
    for depth in depths:
        ctxs = [mas.Context(0) for _ in range(depth)]
        batches = [mas.Batch(c, desc, B) for c in ctxs]
        for b in batches:
            b.set_initial_states(x0)
            for _ in range(3):
                b.set_controls(None)
                b.solve(prm)
        for c in ctxs:
            c.synchronize()
        per = steps_total // depth
        start = threading.Barrier(depth + 1)

        def work(i):
            start.wait()
            time.sleep(i * 0.0169 / depth)
            for _ in range(per):
                batches[i].set_controls(None)
                batches[i].solve(prm)
            ctxs[i].synchronize()

        th = [threading.Thread(target=work, args=(i,)) for i in range(depth)]
        for t in th:
            t.start()
        torch.cuda.synchronize()
        start.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = per * depth
        print(f"depth {depth}: {n} solves of {B} in {dt * 1e3:.1f} ms -> {dt * 1e3 / n:.2f} ms/solve, {B * n / dt / 1e6:.2f} M solves/s", flush=True)
        for b in batches:
            b.close()


if __name__ == "__main__":
    main()
