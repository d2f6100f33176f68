"""Single-solve latency of the 65,536-problem batch with per-kernel event timing, for A/B runs: args = trial_store(0/1)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

store = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
ctx = mas.Context(0)
b = mas.Batch(ctx, mas.example_desc(0), B)
b.set_trial_store(store)
b.set_initial_states(mas.synthetic_single_track_x0(B))
prm = mas.IlqrParams.make(10, 1e-5)
for _ in range(3):
    b.set_controls(None)
    b.solve(prm)
ctx.synchronize()
b.set_profiling(True)
for _ in range(4):
    b.set_controls(None)
    b.solve(prm)
ctx.synchronize()
p = b.profile()
n = max(p["solves"], 1)
print(f"batch {B} trial_store={store}: prologue {p['prologue_ms'] / n:.3f} backward {p['backward_ms'] / n:.3f} forward {p['forward_ms'] / n:.3f} ms per solve; "
      f"total {(p['prologue_ms'] + p['backward_ms'] + p['forward_ms']) / n:.3f}")
