import sys, numpy as np
sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas
ctx = mas.Context(0)
th = 2.0 * np.pi * np.arange(32) / 32
xe = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
d1 = mas.example_desc(1); prm = mas.IlqrParams.make(100, 1e-5)
ref = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, prm, 1, xe)
bad = 0
for rep in range(6):
    r = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, prm, 1, np.repeat(xe, 148, axis=0))
    for k in ("X", "U", "costs", "total_cost"):
        a = np.asarray(r[k]); b = np.asarray(ref[k])
        same = all(np.array_equal(a[s], b[0]) for s in range(148))
        bad += 0 if same else 1
print("total", ref["total_cost"][0], "nondeterministic comparisons:", bad)
