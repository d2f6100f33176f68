"""Does a concurrent D2H stream slow the solve kernels?  Resident solves on `depth` lanes with / without a background copy loop."""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

B = 65536
depth, steps_per_lane = 4, 8
desc = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
prm = mas.IlqrParams.make(10, 1e-5)
x0 = mas.synthetic_single_track_x0(B)
lanes = []
for _ in range(depth):
    s = torch.cuda.Stream()
    ctx = mas.Context(0, s.cuda_stream)
    b = mas.Batch(ctx, desc, B)
    b.set_initial_states(x0)
    lanes.append((s, ctx, b))


def work(i):
    torch.cuda.set_device(0)
    b = lanes[i][2]
    for _ in range(steps_per_lane):
        b.set_controls(None)
        b.solve(prm)
        lanes[i][1].synchronize()


def run(background):
    stop = threading.Event()
    copied = [0]

    def copier():
        torch.cuda.set_device(0)
        cs = torch.cuda.Stream()
        dev = torch.empty(170 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        host = torch.empty(170 * 1024 * 1024, dtype=torch.uint8).pin_memory()
        with torch.cuda.stream(cs):
            while not stop.is_set():
                if background == "d2h":
                    host.copy_(dev, non_blocking=True)
                else:
                    dev.copy_(host, non_blocking=True)
                cs.synchronize()
                copied[0] += 1

    bg = None
    if background:
        bg = threading.Thread(target=copier)
        bg.start()
        time.sleep(0.05)
    th = [threading.Thread(target=work, args=(i,)) for i in range(depth)]
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - w0
    n0 = copied[0]
    stop.set()
    if bg:
        bg.join()
    print(f"background={background}: {wall * 1e3 / (depth * steps_per_lane):.2f} ms/step, copies during run {n0} ({n0 * 0.178 / wall:.1f} GB/s)")


run(None)
run(None)
run("d2h")
run("h2d")
run(None)
