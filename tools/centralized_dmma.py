"""Config 5 (32 stacked circular-track agents, n_s = 128, m_s = 64, all-FD) with the dense gain / value-update products
on the fp64 tensor cores (MAS_B200_CENTRALIZED_DMMA=1, opt-in) against the default bit-exact path: speed-up, and the
deviation of the results next to the reference's own sensitivity to a one-ulp change of the input (oracle, portable trig).

    python tools/centralized_dmma.py > gpurun_out/centralized_dmma.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multi_agent_solver_b200 as mas  # noqa: E402
from oracle import oracle_py as o  # noqa: E402


def best_of(fn, repeats=3):
    fn()
    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def run(ctx, x, dmma, repeats=3):
    os.environ["MAS_B200_CENTRALIZED_DMMA"] = "1" if dmma else "0"
    d1 = mas.example_desc(1)
    p100 = mas.IlqrParams.make(100, 1e-5)
    t = best_of(lambda: mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x, trace=False), repeats)
    r = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, d1, p100, 1, x)
    return t, r


def deviation(a, b):
    return {"rel_total_cost": float(np.max(np.abs(a["total_cost"] - b["total_cost"]) / np.abs(b["total_cost"]))),
            "max_abs_dU": float(np.abs(a["U"] - b["U"]).max()), "max_abs_dX": float(np.abs(a["X"] - b["X"]).max()),
            "iterations": [int(a["trace_iters"][0, 0, 0]), int(b["trace_iters"][0, 0, 0])]}


def main():
    ctx = mas.Context(0)
    th = 2.0 * np.pi * np.arange(32) / 32
    xe = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
    out = {"what": "multi_agent_single_track --agents 32 --strategy centralized; default = k-ascending unfused sums (bit-exact with the "
                   "reference), dmma = mma.sync.m8n8k4.f64 for K = -Q_uu^-1 Q_ux, K^T Q_uu and V_xx (opt-in, not in the parity gate)"}
    t0, r0 = run(ctx, xe, False)
    t1, r1 = run(ctx, xe, True)
    out["one_scenario_ms"] = {"default": t0 * 1e3, "dmma": t1 * 1e3}
    out["deviation_dmma_vs_default"] = deviation(r1, r0)
    it_d, it_0 = out["deviation_dmma_vs_default"]["iterations"]
    out["ms_per_iteration"] = {"default": t0 * 1e3 / it_0, "dmma": t1 * 1e3 / it_d, "speedup": (t0 / it_0) / (t1 / it_d),
                               "note": "the two runs take different numbers of iterations (rounding moves the all-FD problem), so the speed-up is per iteration"}
    x592 = np.repeat(xe, 592, axis=0)
    t0b, _ = run(ctx, x592, False, 2)
    t1b, _ = run(ctx, x592, True, 2)
    out["replicas_592_scenarios_per_s"] = {"default": 592 / t0b, "dmma": 592 / t1b}
    # the reference's own band: the oracle (== the reference's code, tests/test_ref_pin.py) on the same scenario with the
    # first agent's x moved by one ulp
    a = o.strategy_run_batch(o.STRATEGY_CENTRALIZED, o.MODEL_ST_CIRC, xe, trig=o.TRIG_PORTABLE)
    xp = xe.copy()
    xp[0, 0, 0] = np.nextafter(xp[0, 0, 0], np.inf)
    b = o.strategy_run_batch(o.STRATEGY_CENTRALIZED, o.MODEL_ST_CIRC, xp, trig=o.TRIG_PORTABLE)
    out["reference_one_ulp_band"] = deviation(a, b)
    out["default_vs_oracle_bit_equal"] = bool(np.array_equal(r0["U"], a["U"]) and np.array_equal(r0["total_cost"], a["total_cost"]))
    os.environ["MAS_B200_CENTRALIZED_DMMA"] = "0"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
