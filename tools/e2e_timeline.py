"""Host-side timeline of the e2e pipeline (bench.py's e2e_step): how long each lane spends in upload / solve / download."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

B, T, NX, NU = 65536, 80, 4, 2
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps_per_lane = int(sys.argv[2]) if len(sys.argv) > 2 else 8
skip_x = len(sys.argv) > 3 and sys.argv[3] == "noX"
desc = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
prm = mas.IlqrParams.make(10, 1e-5)
x0 = torch.from_numpy(mas.synthetic_single_track_x0(B)).pin_memory().numpy()
lanes = []
for _ in range(depth):
    s = torch.cuda.Stream()
    ctx = mas.Context(0, s.cuda_stream)
    b = mas.Batch(ctx, desc, B)
    out = dict(X=None if skip_x else torch.empty((B, T + 1, NX), dtype=torch.float64).pin_memory().numpy(),
               U=torch.empty((B, T, NU), dtype=torch.float64).pin_memory().numpy(), cost=torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
               iterations=torch.empty(B, dtype=torch.int32).pin_memory().numpy(), status=torch.empty(B, dtype=torch.int32).pin_memory().numpy())
    lanes.append((s, ctx, b, out))
log = [[] for _ in range(depth)]


def work(i, n):
    torch.cuda.set_device(0)
    _, _, b, out = lanes[i]
    for _ in range(n):
        t0 = time.perf_counter()
        b.set_initial_states(x0)
        b.set_controls(None)
        t1 = time.perf_counter()
        b.solve(prm)
        t2 = time.perf_counter()
        b.get_solution(out)
        t3 = time.perf_counter()
        log[i].append((t1 - t0, t2 - t1, t3 - t2))


for rep in range(2):
    for l in log:
        l.clear()
    th = [threading.Thread(target=work, args=(i, steps_per_lane)) for i in range(depth)]
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
a = np.array([x for l in log for x in l[2:]]) * 1e3
print(f"depth {depth} steps/lane {steps_per_lane} skipX {skip_x}: {wall * 1e3 / (depth * steps_per_lane):.2f} ms/step; per lane-step upload {a[:, 0].mean():.2f} "
      f"solve {a[:, 1].mean():.2f} download {a[:, 2].mean():.2f} ms (cycle {a.sum(1).mean():.2f})")
