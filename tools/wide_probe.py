"""One solve alone, small batches: Riccati sweep with one lane per matrix entry (RiccatiWide, MAS_B200_SWEEP_WIDE=1) against
the four-lane sweep (=0) and the default thresholds (unset).  One JSON line per case.

    for w in 0 1 ""; do MAS_B200_SWEEP_WIDE=$w python tools/wide_probe.py; done
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402


def best_of(fn, repeats=7):
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ctx = mas.Context(0)
    x_all = mas.synthetic_single_track_x0(65536)
    p10 = mas.IlqrParams.make(10, 1e-5)
    d0 = mas.example_desc(0)
    wide = os.environ.get("MAS_B200_SWEEP_WIDE", "auto") or "auto"
    for B in (1, 256, 1024, 2048, 4096, 8192, 65536):
        x0 = np.array([[0.0, 1.0, 0.0, 0.0]]) if B == 1 else x_all[:B]
        b = mas.Batch(ctx, d0, B)
        b.set_initial_states(x0)

        def solve():
            b.set_controls(None)
            b.solve(p10)
            ctx.synchronize()

        t = best_of(solve)
        print(json.dumps({"case": f"ST-lane x {B}, resident, one solve", "sweep_wide": wide, "ms": round(t * 1e3, 4)}), flush=True)
        b.close()


if __name__ == "__main__":
    main()
