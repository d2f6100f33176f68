"""One single-track problem (config 1), a few solves: the target of ncu captures of the latency path."""
import sys

import numpy as np

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

ctx = mas.Context(0)
b = mas.Batch(ctx, mas.example_desc(0), 1)
b.set_initial_states(np.array([[0.0, 1.0, 0.0, 0.0]]))
prm = mas.IlqrParams.make(10, 1e-5)
for _ in range(3):
    b.set_controls(None)
    b.solve(prm)
ctx.synchronize()
print(b.get_solution()["cost"])
