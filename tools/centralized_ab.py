"""Config 5 timing of the library selected by MAS_B200_LIB (A/B of kernel variants): one scenario, 592 replicas, checksum."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas
from bench import best_of
ctx = mas.Context(0)
th = 2.0 * np.pi * np.arange(32) / 32
xe = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
d1 = mas.example_desc(1); p100 = mas.IlqrParams.make(100, 1e-5)
t1 = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, xe, trace=False), 3)
x5 = np.repeat(xe, 592, axis=0)
t5 = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x5, trace=False), 2)
r = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, xe)
print(json.dumps({"lib": os.environ.get("MAS_B200_LIB", "default"), "one_scenario_ms": t1 * 1e3, "replicas_592_per_s": 592 / t5, "total_cost": float(r["total_cost"][0]), "iterations": int(r["trace_iters"][0,0,0])}))
