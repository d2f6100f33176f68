"""Generates tests/golden/*.npz.

PROVENANCE: the reference (markomiz/multi_agent_solver) cannot be built or run in this image (it needs
Eigen 3.4; no network), and its repository holds no golden vectors for iLQR.  These fixtures are therefore
REGRESSION PINS produced by oracle/ (the CPU restatement of the reference), not reference outputs:
they freeze today's oracle so that later edits to the oracle or to the shared portable trig cannot drift
unnoticed, and they let the GPU tests compare against committed numbers.  Both libm modes are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle_py as o  # noqa: E402


def config3_x0(n):
    """First n problems of the config-3 batch: std::mt19937_64(20240607), Y, psi, v."""
    import multi_agent_solver_b200 as mas

    return mas.synthetic_single_track_x0(n)


def main():
    out = {}
    # config 1: single_track_ocp
    for trig, tag in ((o.TRIG_GLIBC, "glibc"), (o.TRIG_PORTABLE, "portable")):
        r = o.ilqr_solve_batch(o.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iterations=10, tolerance=1e-5, trig=trig)
        for k in ("X", "U", "cost", "iterations", "status"):
            out[f"config1_{tag}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "config1_single_track.npz"), **out)

    # config 3: first 96 problems of the headline batch
    x0 = config3_x0(96)
    out = {"x0": x0}
    for trig, tag in ((o.TRIG_GLIBC, "glibc"), (o.TRIG_PORTABLE, "portable")):
        r = o.ilqr_solve_batch(o.MODEL_ST_LANE, x0, max_iterations=10, tolerance=1e-5, trig=trig)
        for k in ("cost", "iterations", "status", "alpha_trials"):
            out[f"{tag}_{k}"] = r[k]
        out[f"{tag}_U_final_step0"] = r["U"][:, 0, :]
        out[f"{tag}_X_terminal"] = r["X"][:, -1, :]
    np.savez_compressed(os.path.join(HERE, "config3_first96.npz"), **out)

    # config 2: three agents, trust region, 10 outer rounds
    th = 2.0 * np.pi * np.arange(3) / 3
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(3, 4.0)], -1)[None]
    out = {"x0": x0}
    for trig, tag in ((o.TRIG_GLIBC, "glibc"), (o.TRIG_PORTABLE, "portable")):
        r = o.strategy_run_batch(o.STRATEGY_TRUSTREGION, o.MODEL_ST_CIRC, x0, max_outer=10, max_iterations=100, tolerance=1e-5, trig=trig)
        for k in ("X", "U", "costs", "total_cost", "trace_iters", "trace_accept", "trace_cost"):
            out[f"{tag}_{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "config2_trust_region_3agents.npz"), **out)

    # config 4: LQR agents, sequential, 10 outer rounds (all agents identical: 4 are enough)
    x0 = np.tile([1.0, 0.0, 0.0, 0.0], (1, 4, 1))
    r = o.strategy_run_batch(o.STRATEGY_SEQUENTIAL, o.MODEL_LQR, x0, max_outer=10, max_iterations=100, tolerance=1e-5, trig=o.TRIG_PORTABLE)
    np.savez_compressed(os.path.join(HERE, "config4_sequential_lqr.npz"), x0=x0, **{k: r[k] for k in ("X", "U", "costs", "total_cost", "trace_iters")})

    # config 5 (small): centralized, 4 stacked circular-track agents
    th = 2.0 * np.pi * np.arange(4) / 4
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(4, 4.0)], -1)[None]
    r = o.strategy_run_batch(o.STRATEGY_CENTRALIZED, o.MODEL_ST_CIRC, x0, max_outer=1, max_iterations=100, tolerance=1e-5, trig=o.TRIG_PORTABLE)
    np.savez_compressed(os.path.join(HERE, "config5_centralized_4agents.npz"), x0=x0, iterations=r["trace_iters"][:, 0, 0],
                        **{k: r[k] for k in ("X", "U", "costs", "total_cost")})
    # config 5 at full size: 32 stacked circular-track agents (n = 128, m = 64); ~10 s in the oracle
    th = 2.0 * np.pi * np.arange(32) / 32
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(32, 4.0)], -1)[None]
    r = o.strategy_run_batch(o.STRATEGY_CENTRALIZED, o.MODEL_ST_CIRC, x0, max_outer=1, max_iterations=100, tolerance=1e-5, trig=o.TRIG_PORTABLE)
    np.savez_compressed(os.path.join(HERE, "config5_centralized_32agents.npz"), x0=x0, iterations=r["trace_iters"][:, 0, 0],
                        **{k: r[k] for k in ("X", "U", "costs", "total_cost")})
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
