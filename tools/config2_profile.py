import sys, numpy as np, time
sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas
sys.path.insert(0, "tools")
from bench_configs import circle
ctx = mas.Context(0)
x0, gp, op = circle(4096, 3)
B = 4096 * 3
d1 = mas.example_desc(1)
b = mas.Batch(ctx, d1, B)
b.set_initial_states(x0.reshape(B, 4)); b.set_params(gp.reshape(B, 6)); b.set_controls(None)
prm = mas.IlqrParams.make(100, 1e-5)
b.solve(prm); ctx.synchronize()
b.set_profiling(True)
for _ in range(3):
    b.set_controls(None); b.solve(prm)
ctx.synchronize()
p = b.profile(); n = p["solves"]
st = b.stats()
print({k: (round(v / n, 3) if isinstance(v, float) else v) for k, v in p.items()}, st["outer_iterations_run"], st["forward_lanes"])
t0 = time.perf_counter()
b.set_profiling(False)
for _ in range(3):
    b.set_controls(None); b.solve(prm)
ctx.synchronize()
print("solve ms", (time.perf_counter() - t0) / 3 * 1e3)
