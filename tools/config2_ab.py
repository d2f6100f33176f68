"""Config 2 (4,096 scenarios x 3 all-FD circular-track agents, trust region, 10 rounds) under MAS_B200_BACKWARD_MODE."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402
from bench import circle_scenarios, best_of  # noqa: E402

ctx = mas.Context(0)
S, A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 3
x0, gp = circle_scenarios(S, A)
d1 = mas.example_desc(1)
p100 = mas.IlqrParams.make(100, 1e-5)
t = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, d1, p100, 10, x0, model_params=gp, trace=False), 4)
r = mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, d1, p100, 10, x0, model_params=gp)
print(json.dumps({"mode": os.environ.get("MAS_B200_BACKWARD_MODE", "0"), "scenarios": S, "ms": t * 1e3, "scenarios_per_s": S / t,
                  "iters_total": int(r["trace_iters"].sum()), "checksum": float(r["total_cost"].sum())}))
