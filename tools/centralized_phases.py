"""Config 5 (32 stacked single-track agents, centralized): where one scenario's cycles go, alone and with every SM busy.

    MAS_B200_CENTRALIZED_PHASES=1 python tools/centralized_phases.py [replicas]
"""
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
import multi_agent_solver_b200 as mas  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = mas.Context(0)
th = 2.0 * np.pi * np.arange(32) / 32
xe = np.repeat(np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None], S, axis=0)
d1 = mas.example_desc(1)
prm = mas.IlqrParams.make(100, 1e-5)
mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, prm, 1, xe, trace=False)
clocks = []
stop = threading.Event()


def sample():
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True,
                                 text=True, timeout=5).stdout.strip().splitlines()[0]
            clocks.append(out)
        except Exception:
            pass


t = threading.Thread(target=sample)
t.start()
t0 = time.perf_counter()
for _ in range(8):
    r = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, prm, 1, xe, trace=False)
dt = (time.perf_counter() - t0) / 8
stop.set()
t.join()
print(f"{S} replicas: {dt * 1e3:.1f} ms per call, total cost {r['total_cost'][0]:.10f}; clocks/power under load: {clocks[-3:]}")
