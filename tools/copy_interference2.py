"""Which part of a solve interferes with a concurrent D2H stream?  4 lanes run a load, one thread copies D2H."""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

B = 65536
depth = 4
desc = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
x0 = mas.synthetic_single_track_x0(B)
lanes = []
for _ in range(depth):
    s = torch.cuda.Stream()
    ctx = mas.Context(0, s.cuda_stream)
    b = mas.Batch(ctx, desc, B)
    b.set_initial_states(x0)
    lanes.append((s, ctx, b))
n = 170 * 1024 * 1024
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()


def run(kind, background, seconds=0.6):
    stop = threading.Event()
    done = [0] * depth
    copied = [0]

    def work(i):
        torch.cuda.set_device(0)
        _, ctx, b = lanes[i]
        prm = mas.IlqrParams.make(1 if kind == "solve1" else 10, 1e-5)
        while not stop.is_set():
            if kind == "prologue":
                for _ in range(8):
                    b.initialize()
            else:
                b.set_controls(None)
                b.solve(prm)
            ctx.synchronize()
            done[i] += 1

    def copier():
        torch.cuda.set_device(0)
        cs = torch.cuda.Stream()
        with torch.cuda.stream(cs):
            while not stop.is_set():
                host.copy_(dev, non_blocking=True)
                cs.synchronize()
                copied[0] += 1

    th = [threading.Thread(target=work, args=(i,)) for i in range(depth)]
    if background:
        th.append(threading.Thread(target=copier))
    for t in th:
        t.start()
    time.sleep(0.2)
    d0, c0, t0 = sum(done), copied[0], time.perf_counter()
    time.sleep(seconds)
    d1, c1, t1 = sum(done), copied[0], time.perf_counter()
    stop.set()
    for t in th:
        t.join()
    print(f"{kind:9s} background={background!s:5}: {(d1 - d0) / (t1 - t0):8.1f} load units/s, D2H {(c1 - c0) * n / (t1 - t0) / 1e9:5.1f} GB/s", flush=True)


for kind in ("prologue", "solve1", "solve"):
    run(kind, False)
    run(kind, True)
