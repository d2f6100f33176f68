"""The committed fixtures in tests/golden/ (tests/golden/make_golden.py) against the oracle, against a build of the
reference's own sources (oracle/_ref/libref.so: test_fixtures_are_outputs_of_the_reference_build) and, in the gpu tests,
against the CUDA path with no checker involved at run time.  The portable-trig fixtures must reproduce bit for bit; the
glibc ones to 1e-9 / 1e-7 on the problems that are not ill-conditioned (libm may differ between hosts in the last bit).
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def test_config1(oracle):
    g = load("config1_single_track.npz")
    r = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "cost", "iterations", "status"):
        assert np.array_equal(r[k], g[f"config1_portable_{k}"]), k
    r = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_GLIBC)
    assert abs(r["cost"][0] - g["config1_glibc_cost"][0]) <= 1e-9 * abs(g["config1_glibc_cost"][0])
    assert np.max(np.abs(r["X"] - g["config1_glibc_X"])) <= 1e-7
    assert np.array_equal(r["iterations"], g["config1_glibc_iterations"])
    # the two libm modes agree on this well-conditioned problem far below the parity tolerance
    assert abs(g["config1_glibc_cost"][0] - g["config1_portable_cost"][0]) <= 1e-12 * g["config1_glibc_cost"][0]


def test_config3_first96(oracle, mas):
    g = load("config3_first96.npz")
    assert np.array_equal(mas.synthetic_single_track_x0(96), g["x0"])
    r = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, g["x0"], max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert np.array_equal(r["cost"], g["portable_cost"])
    assert np.array_equal(r["iterations"], g["portable_iterations"]) and np.array_equal(r["status"], g["portable_status"])
    assert np.array_equal(r["alpha_trials"], g["portable_alpha_trials"])
    assert np.array_equal(r["U"][:, 0, :], g["portable_U_final_step0"]) and np.array_equal(r["X"][:, -1, :], g["portable_X_terminal"])
    assert np.array_equal(g["glibc_iterations"], g["portable_iterations"])


def test_config2_trust_region(oracle):
    g = load("config2_trust_region_3agents.npz")
    r = oracle.strategy_run_batch(oracle.STRATEGY_TRUSTREGION, oracle.MODEL_ST_CIRC, g["x0"], max_outer=10, max_iterations=100, tolerance=1e-5,
                                  trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost", "trace_iters", "trace_accept", "trace_cost"):
        assert np.array_equal(r[k], g[f"portable_{k}"]), k
    assert np.array_equal(g["glibc_trace_iters"], g["portable_trace_iters"])


def test_config4_sequential(oracle):
    g = load("config4_sequential_lqr.npz")
    r = oracle.strategy_run_batch(oracle.STRATEGY_SEQUENTIAL, oracle.MODEL_LQR, g["x0"], max_outer=10, max_iterations=100, tolerance=1e-5,
                                  trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost", "trace_iters"):
        assert np.array_equal(r[k], g[k]), k


def test_config5_centralized_small(oracle):
    g = load("config5_centralized_4agents.npz")
    r = oracle.strategy_run_batch(oracle.STRATEGY_CENTRALIZED, oracle.MODEL_ST_CIRC, g["x0"], max_outer=1, max_iterations=100, tolerance=1e-5,
                                  trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(r[k], g[k]), k
    assert r["trace_iters"][0, 0, 0] == g["iterations"][0]


def test_fixtures_are_outputs_of_the_reference_build():
    """Every portable-trig fixture equals, bit for bit, what the reference's OWN code returns for the fixture's inputs
    (oracle/_ref/libref.so: /root/reference's unmodified headers and example OCP builders on oracle/eigen_shim).  The
    32-agent config-5 fixture is covered by tests/test_ref_pin.py::test_config5_centralized_32_agents (same inputs)."""
    from oracle import ref_py as ref

    if not ref.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference sources to build it from")
    ref.build()
    g = load("config1_single_track.npz")
    for trig, tag in ((ref.TRIG_GLIBC, "glibc"), (ref.TRIG_PORTABLE, "portable")):
        r = ref.ilqr_solve_batch(ref.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iterations=10, tolerance=1e-5, trig=trig)
        if trig == ref.TRIG_PORTABLE:
            for k in ("X", "U", "cost", "iterations", "status"):
                assert np.array_equal(r[k], g[f"config1_{tag}_{k}"]), k
        else:  # the host's libm may differ from the generating host's in the last bit
            assert abs(r["cost"][0] - g["config1_glibc_cost"][0]) <= 1e-9 * abs(g["config1_glibc_cost"][0])
    g = load("config3_first96.npz")
    assert np.array_equal(ref.synthetic_single_track_x0(96), g["x0"])
    r = ref.ilqr_solve_batch(ref.MODEL_ST_LANE, g["x0"], max_iterations=10, tolerance=1e-5, trig=ref.TRIG_PORTABLE)
    for k in ("cost", "iterations", "status", "alpha_trials"):
        assert np.array_equal(r[k], g[f"portable_{k}"]), k
    assert np.array_equal(r["U"][:, 0, :], g["portable_U_final_step0"]) and np.array_equal(r["X"][:, -1, :], g["portable_X_terminal"])
    g = load("config2_trust_region_3agents.npz")
    r = ref.strategy_run_batch(ref.STRATEGY_TRUSTREGION, ref.MODEL_ST_CIRC, g["x0"], max_outer=10, max_iterations=100, tolerance=1e-5,
                               trig=ref.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(r[k], g[f"portable_{k}"]), k
    assert np.array_equal(r["iterations_total"], g["portable_trace_iters"].sum(1))
    g = load("config4_sequential_lqr.npz")
    r = ref.strategy_run_batch(ref.STRATEGY_SEQUENTIAL, ref.MODEL_LQR, g["x0"], max_outer=10, max_iterations=100, tolerance=1e-5,
                               trig=ref.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(r[k], g[k]), k
    assert np.array_equal(r["iterations_total"], g["trace_iters"].sum(1))
    g = load("config5_centralized_4agents.npz")
    r = ref.strategy_run_batch(ref.STRATEGY_CENTRALIZED, ref.MODEL_ST_CIRC, g["x0"], max_outer=1, max_iterations=100, tolerance=1e-5,
                               trig=ref.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(r[k], g[k]), k
    assert r["iterations_total"][0, 0] == g["iterations"][0]


def test_config5_centralized_32_agents_device_source(emu):
    """BASELINE configs[4] at full size (32 stacked agents, n = 128, m = 64): the device source, emulated on
    the host, against the committed oracle output (the oracle itself needs ~10 s for this case)."""
    g = load("config5_centralized_32agents.npz")
    got = emu.solve_centralized(1, g["x0"][0])
    assert got["iterations"] == g["iterations"][0] and got["total_cost"] == g["total_cost"][0]
    assert np.array_equal(got["X"], g["X"][0]) and np.array_equal(got["U"], g["U"][0]) and np.array_equal(got["costs"], g["costs"][0])


@pytest.mark.gpu
def test_gpu_centralized_matches_golden(mas, ctx):
    """Centralized strategy on the GPU: 32-agent stacked problem (config 5) and the 4-agent one."""
    for name in ("config5_centralized_32agents.npz", "config5_centralized_4agents.npz"):
        g = load(name)
        x0 = np.repeat(g["x0"], 3, axis=0)  # three identical scenarios: one CTA each
        r = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 1, x0)
        for s in range(3):
            assert r["trace_iters"][s, 0, 0] == g["iterations"][0]
            assert r["total_cost"][s] == g["total_cost"][0]
            assert np.array_equal(r["X"][s], g["X"][0]) and np.array_equal(r["U"][s], g["U"][0]) and np.array_equal(r["costs"][s], g["costs"][0])


@pytest.mark.gpu
def test_gpu_matches_golden(mas, ctx):
    """The CUDA path against the committed portable-trig fixtures (no oracle involved at run time)."""
    g = load("config3_first96.npz")
    got = mas.ilqr_solve_batch(ctx, mas.example_desc(0), mas.IlqrParams.make(10, 1e-5), g["x0"])
    assert np.array_equal(got["cost"], g["portable_cost"])
    assert np.array_equal(got["iterations"], g["portable_iterations"]) and np.array_equal(got["status"], g["portable_status"])
    assert np.array_equal(got["U"][:, 0, :], g["portable_U_final_step0"]) and np.array_equal(got["X"][:, -1, :], g["portable_X_terminal"])
    g1 = load("config1_single_track.npz")
    got = mas.ilqr_solve_batch(ctx, mas.example_desc(0), mas.IlqrParams.make(10, 1e-5), np.array([[0.0, 1.0, 0.0, 0.0]]))
    assert np.array_equal(got["X"], g1["config1_portable_X"]) and np.array_equal(got["U"], g1["config1_portable_U"])
    g2 = load("config2_trust_region_3agents.npz")
    r = mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 10, g2["x0"])
    for k in ("X", "U", "costs", "total_cost", "trace_iters", "trace_accept", "trace_cost"):
        assert np.array_equal(r[k], g2[f"portable_{k}"]), k
    g4 = load("config4_sequential_lqr.npz")
    r = mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, mas.example_desc(2), mas.IlqrParams.make(100, 1e-5), 10, g4["x0"])
    for k in ("X", "U", "costs", "total_cost", "trace_iters"):
        assert np.array_equal(r[k], g4[k]), k
