"""The oracle (oracle/, CPU restatement of the reference) against the known answers that do not need the reference
build: closed-form anchors, the reference's unit-test values for the layers under iLQR (oracle_selftest.cpp follows
tests/ocp_tests.cpp:21-154), and the SURVEY section 9 probe values (an independent numpy restatement).  The pin to the
reference's own executed code is tests/test_ref_pin.py (oracle == oracle/_ref/libref.so bit for bit).
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_unit_test_values(oracle):
    """tests/ocp_tests.cpp:21-154 of the reference: shapes, best_cost == 0, FD defaults installed,
    id-sorted blocks, concatenated bounds, exact stacked dynamics / stage / terminal values."""
    exe = os.path.join(ROOT, "oracle", "_build", "oracle_selftest")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout


def test_initial_cost_anchor_single_track(oracle):
    """J(U=0) = 80 * (10*1^2 + 1*(0-1)^2) = 880 exactly for x0 = (0,1,0,0) (car at rest)."""
    X, cost = oracle.rollout_cost(oracle.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), np.zeros((1, 80, 2)))
    assert cost[0] == 880.0
    assert np.array_equal(X[0], np.tile([0.0, 1.0, 0.0, 0.0], (81, 1)))


def test_initial_cost_anchor_lqr(oracle):
    """LQR with U=0: x_{t+1} = x_t (1 + dt + dt^2/2 + dt^3/6 + dt^4/24) per RK4 step on xdot = x."""
    g = 1 + 0.1 + 0.1**2 / 2 + 0.1**3 / 6 + 0.1**4 / 24
    expect = sum(g ** (2 * t) for t in range(10)) + g**20
    _, cost = oracle.rollout_cost(oracle.MODEL_LQR, np.array([[1.0, 0, 0, 0]]), np.zeros((1, 10, 4)))
    assert abs(cost[0] - expect) < 1e-12 * expect


def test_config1_trace(oracle):
    """single_track_ocp, 10 / 1e-5: SURVEY 9 P2 -- 823.2799586777 (alpha 0.25), 508.5930603049 (alpha
    0.0625), third line search fails -> stop; 3 backward passes, 18 trial rollouts + the prologue."""
    for trig in (oracle.TRIG_GLIBC, oracle.TRIG_PORTABLE):
        r = oracle.ilqr_solve_trace(oracle.MODEL_ST_LANE, [0, 1, 0, 0], max_iterations=10, tolerance=1e-5, trig=trig)
        assert len(r["cost_trace"]) == 3
        assert abs(r["cost_trace"][0] - 823.2799586777) < 1e-9
        assert abs(r["cost_trace"][1] - 508.5930603049) < 1e-9
        assert r["cost_trace"][2] == r["cost_trace"][1]
        assert list(r["alpha_index"]) == [2, 4, -1]
        assert r["alpha_trials"] == 18 and r["rollouts"] == 19 and r["reg_retries"] == 0
    b = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, np.array([[0.0, 1, 0, 0]]), max_iterations=10, tolerance=1e-5)
    assert b["iterations"][0] == 3 and b["status"][0] == oracle.STATUS_CONVERGED


def test_lqr_agent(oracle):
    """multi_agent_lqr agent: SURVEY 9 P5 -- 20.869847032359 in 3 iterations / 13 trial rollouts; a
    second solve from the solution is one iteration of ten failed candidates."""
    r = oracle.ilqr_solve_batch(oracle.MODEL_LQR, np.array([[1.0, 0, 0, 0]]), max_iterations=100, tolerance=1e-5)
    assert abs(r["cost"][0] - 20.869847032359) < 1e-11
    assert r["iterations"][0] == 3 and r["alpha_trials"][0] == 13
    again = oracle.ilqr_solve_batch(oracle.MODEL_LQR, np.array([[1.0, 0, 0, 0]]), U_init=r["U"], max_iterations=100, tolerance=1e-5)
    assert again["iterations"][0] == 1 and again["alpha_trials"][0] == 10 and again["cost"][0] == r["cost"][0]


def test_trust_region_three_agents(oracle):
    """multi_agent_single_track --agents 3 --strategy trustregion (config 2): SURVEY 9 P6 -- per agent
    round 0 accepted after 4 iterations (18 Q_uu regularisation retries), round 1 accepted after 8,
    rounds 2-9 rejected with 1 iteration each; costs ~25.9 then ~2.0445 (chaos-sensitive digits)."""
    th = 2.0 * np.pi * np.arange(3) / 3
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(3, 4.0)], -1)[None]
    r = oracle.strategy_run_batch(oracle.STRATEGY_TRUSTREGION, oracle.MODEL_ST_CIRC, x0, max_outer=10, max_iterations=100, tolerance=1e-5)
    assert np.array_equal(r["trace_iters"][0], np.array([[4] * 3, [8] * 3] + [[1] * 3] * 8))
    assert np.array_equal(r["trace_accept"][0], np.array([[1] * 3, [1] * 3] + [[0] * 3] * 8))
    assert np.all(np.abs(r["trace_cost"][0][0] - 25.88) < 0.05)
    assert np.all(np.abs(r["costs"][0] - 2.04452) < 5e-5)
    assert abs(r["total_cost"][0] - r["costs"][0].sum()) < 1e-12
    first = oracle.ilqr_solve_batch(oracle.MODEL_ST_CIRC, x0[0], max_iterations=100, tolerance=1e-5)
    assert list(first["reg_retries"]) == [18, 18, 18]


def test_sequential_lqr(oracle):
    """multi_agent_lqr, sequential (config 4): exactly max_outer rounds; 3 iterations then 1 per round."""
    x0 = np.tile([1.0, 0, 0, 0], (1, 5, 1))
    r = oracle.strategy_run_batch(oracle.STRATEGY_SEQUENTIAL, oracle.MODEL_LQR, x0, max_outer=4, max_iterations=100, tolerance=1e-5)
    assert np.array_equal(r["trace_iters"][0], np.array([[3] * 5] + [[1] * 5] * 3))
    assert np.all(np.abs(r["costs"] - 20.869847032359) < 1e-11)


def test_centralized_stacks_agents(oracle):
    """centralized strategy on 3 circular-track agents: SURVEY 9 P7 -- 4 iterations, cost ~18.07."""
    th = 2.0 * np.pi * np.arange(3) / 3
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(3, 4.0)], -1)[None]
    r = oracle.strategy_run_batch(oracle.STRATEGY_CENTRALIZED, oracle.MODEL_ST_CIRC, x0, max_outer=1, max_iterations=100, tolerance=1e-5)
    assert r["trace_iters"][0, 0, 0] == 4
    assert abs(r["total_cost"][0] - 18.07) < 0.05
    # stacked evaluation: block-diagonal dynamics, costs summed in id order
    Xs = np.arange(12, dtype=float) * 0.1 + 1.0
    Us = np.linspace(-0.3, 0.3, 6)
    dyn, stage, term, dims = oracle.global_ocp_eval(oracle.MODEL_ST_CIRC, x0[0], Xs, Us)
    assert list(dims) == [12, 6, 10]
    assert term == 0.0
    per_agent = 0.0
    for a in range(3):
        x, u = Xs[4 * a:4 * a + 4], Us[2 * a:2 * a + 2]
        d = abs(np.sqrt(x[0] * x[0] + x[1] * x[1]) - 20.0)
        per_agent += 1.0 * d * d + 1.0 * (x[3] - 5.0) * (x[3] - 5.0) + 0.001 * u[0] * u[0] + 0.001 * u[1] * u[1]
        assert dyn[4 * a + 3] == u[1]
    assert abs(stage - per_agent) < 1e-12 * per_agent


def test_reference_is_ill_conditioned_on_fd_configs(oracle):
    """Why parity is asserted against the portable-trig oracle: with the reference's own glibc trig, a
    one-ulp change of x0 already moves a large share of config-3 problems by more than the 1e-9 / 1e-7
    tolerances (FD cross-term noise, finite_differences.hpp:263-287), while iteration counts hold."""
    rng = np.random.default_rng(7)
    B = 400
    x0 = np.stack([np.zeros(B), rng.uniform(-2, 2, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 2, B)], -1)
    a = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, trig=oracle.TRIG_GLIBC)
    x0p = x0.copy()
    x0p[:, 1] = np.nextafter(x0p[:, 1], np.inf)
    c = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0p, trig=oracle.TRIG_GLIBC)
    rel = np.abs(a["cost"] - c["cost"]) / np.abs(a["cost"])
    assert (rel > 1e-9).mean() > 0.1
    b = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, trig=oracle.TRIG_PORTABLE)
    assert np.array_equal(a["iterations"], b["iterations"]) and np.array_equal(a["status"], b["status"])
    # the two libm modes stay within the reference's own one-ulp sensitivity band
    rel_modes = np.abs(a["cost"] - b["cost"]) / np.abs(a["cost"])
    assert np.median(rel_modes) < 1e-12 and rel_modes.max() < 5e-2


def test_aliased_symmetrisation_is_benign(oracle):
    """aliased vs exact V_xx symmetrisation give identical bits on configs 1-4."""
    x0 = np.array([[0.0, 1, 0, 0], [0.0, -1.3, 0.2, 1.5]])
    a = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, aliased_sym=True)
    b = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, aliased_sym=False)
    assert np.array_equal(a["X"], b["X"]) and np.array_equal(a["cost"], b["cost"])
