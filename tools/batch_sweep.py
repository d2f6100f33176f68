"""Batch-size sweep of the headline workload (single-track lane following, 10 / 1e-5): time of one solve alone and
throughput with several solves in flight, from the latency end (1,024 problems) to beyond the headline size.
    python tools/batch_sweep.py > gpurun_out/batch_sweep.jsonl
"""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402


def run(batch_size, depth, steps):
    x0 = mas.synthetic_single_track_x0(batch_size)
    desc = mas.example_desc(0)
    prm = mas.IlqrParams.make(10, 1e-5)
    ctxs = [mas.Context(0) for _ in range(depth)]
    batches = [mas.Batch(c, desc, batch_size) for c in ctxs]
    for b in batches:
        b.set_initial_states(x0)
        for _ in range(3):
            b.set_controls(None)
            b.solve(prm)
    for c in ctxs:
        c.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        batches[0].set_controls(None)
        batches[0].solve(prm)
    ctxs[0].synchronize()
    single = (time.perf_counter() - t0) / steps
    go = threading.Barrier(depth + 1)

    def work(i):
        go.wait()
        time.sleep(i * single / depth)
        for _ in range(steps):
            batches[i].set_controls(None)
            batches[i].solve(prm)
        ctxs[i].synchronize()

    th = [threading.Thread(target=work, args=(i,)) for i in range(depth)]
    for t in th:
        t.start()
    torch.cuda.synchronize()
    go.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    piped = (time.perf_counter() - t0) / (steps * depth)
    st = batches[0].stats()
    for b in batches:
        b.close()
    return {"problems": batch_size, "single_solve_ms": single * 1e3, "single_solves_per_s": batch_size / single, "in_flight": depth,
            "pipelined_ms_per_solve": piped * 1e3, "pipelined_solves_per_s": batch_size / piped,
            "mean_iterations": st["iterations"] / batch_size}


if __name__ == "__main__":
    for n, depth in ((1024, 8), (8192, 8), (65536, 4), (262144, 4), (1048576, 2)):
        print(json.dumps(run(n, depth, 6)), flush=True)
