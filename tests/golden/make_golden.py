"""Generates tests/golden/*.npz.

PROVENANCE: the fixtures are written from oracle/ (the CPU restatement of the reference), and every one of them is an
output of the reference's OWN code as well: oracle/_ref/libref.so -- /root/reference's unmodified headers and example OCP
builders compiled against oracle/eigen_shim (the image has no Eigen) -- returns the same bits for the same inputs.
main() asserts that before it writes anything (where libref.so is available), tests/test_golden.py::
test_fixtures_are_outputs_of_the_reference_build re-checks the committed files, and tests/test_ref_pin.py holds the
wider oracle == reference-build comparison.  The reference repository itself holds no golden vectors for iLQR.  Both
libm modes are stored; only the portable-trig arrays are compared bit for bit across hosts.

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


def check_against_reference_build():
    """oracle == reference build on the fixtures' inputs (the strongest statement available offline)."""
    from oracle import ref_py as ref

    if not ref.available():
        print("oracle/_ref/libref.so not available: fixtures written from the oracle alone")
        return
    ref.build()
    x1 = np.array([[0.0, 1.0, 0.0, 0.0]])
    x96 = config3_x0(96)
    for trig in (o.TRIG_GLIBC, o.TRIG_PORTABLE):
        for x0 in (x1, x96):
            a = o.ilqr_solve_batch(o.MODEL_ST_LANE, x0, max_iterations=10, tolerance=1e-5, trig=trig)
            b = ref.ilqr_solve_batch(ref.MODEL_ST_LANE, x0, max_iterations=10, tolerance=1e-5, trig=trig)
            assert all(np.array_equal(a[k], b[k]) for k in ("X", "U", "cost", "iterations", "status", "alpha_trials"))
    th = 2.0 * np.pi * np.arange(3) / 3
    cases = [(o.STRATEGY_TRUSTREGION, o.MODEL_ST_CIRC, np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(3, 4.0)], -1)[None], 10),
             (o.STRATEGY_SEQUENTIAL, o.MODEL_LQR, np.tile([1.0, 0.0, 0.0, 0.0], (1, 4, 1)), 10)]
    for A in (4, 32):
        th = 2.0 * np.pi * np.arange(A) / A
        cases.append((o.STRATEGY_CENTRALIZED, o.MODEL_ST_CIRC, np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(A, 4.0)], -1)[None], 1))
    for kind, model, x0, outer in cases:
        a = o.strategy_run_batch(kind, model, x0, max_outer=outer, max_iterations=100, tolerance=1e-5, trig=o.TRIG_PORTABLE)
        b = ref.strategy_run_batch(kind, model, x0, max_outer=outer, max_iterations=100, tolerance=1e-5, trig=ref.TRIG_PORTABLE)
        assert all(np.array_equal(a[k], b[k]) for k in ("X", "U", "costs", "total_cost"))
        assert np.array_equal(a["trace_iters"].sum(1), b["iterations_total"])
    print("oracle == reference build (oracle/_ref/libref.so) on every fixture input")


def main():
    check_against_reference_build()
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
