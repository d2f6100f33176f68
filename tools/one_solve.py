"""One resident solve of B ST-lane problems with a given backward mode (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402

B, mode = int(sys.argv[1]), int(sys.argv[2])
ctx = mas.Context(0)
x0 = np.array([[0.0, 1.0, 0.0, 0.0]]) if B == 1 else mas.synthetic_single_track_x0(65536)[:B]
b = mas.Batch(ctx, mas.example_desc(0), B)
b.set_backward_mode(mode)
b.set_initial_states(x0)
for _ in range(3):
    b.set_controls(None)
    b.solve(mas.IlqrParams.make(10, 1e-5))
    ctx.synchronize()
print(b.get_solution()["iterations"][:4])
