"""GPU: the multi-agent strategy layer (mas::solve(Strategy&, MultiAgentProblem&)) through the C ABI
against the oracle, round by round: inner iteration counts, accept/reject flags, per-agent costs, final
trajectories and totals.  Oracle in portable-trig mode; bit-identical results required.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def circ_x0(n_scenarios, n_agents, jitter_seed=None):
    th = 2.0 * np.pi * np.arange(n_agents) / n_agents
    R = np.full((n_scenarios, 1), 20.0)
    if jitter_seed is not None:
        R = np.random.default_rng(jitter_seed).uniform(15, 25, (n_scenarios, 1))
    x0 = np.stack([R * np.cos(th), R * np.sin(th), np.broadcast_to(1.57 + th, (n_scenarios, n_agents)), np.full((n_scenarios, n_agents), 4.0)], -1)
    return x0, R


def test_trust_region_config2(mas, ctx, oracle):
    """multi_agent_single_track --agents 3 --strategy trustregion, replicated scenarios with a jittered
    track radius per scenario (BASELINE configs[1])."""
    S, A = 24, 3
    x0, R = circ_x0(S, A, jitter_seed=4)
    op = np.broadcast_to(np.stack([R[:, 0], np.full(S, 5.0)], -1)[:, None, :], (S, A, 2)).copy()
    gp = np.broadcast_to(np.stack([R[:, 0], np.full(S, 5.0), np.ones(S), np.ones(S), np.full(S, 1e-3), np.full(S, 1e-3)], -1)[:, None, :],
                         (S, A, 6)).copy()
    ref = oracle.strategy_run_batch(oracle.STRATEGY_TRUSTREGION, oracle.MODEL_ST_CIRC, x0, params=op, max_outer=10, max_iterations=100,
                                    tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = mas.strategy_run(ctx, mas.Strategy.TRUSTREGION, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 10, x0, model_params=gp)
    for k in ("trace_iters", "trace_accept", "trace_cost", "costs", "total_cost", "X", "U"):
        assert np.array_equal(got[k], ref[k]), k
    assert (ref["trace_accept"] == 0).any() and (ref["trace_accept"] == 1).any()


def test_sequential_config4(mas, ctx, oracle):
    """multi_agent_lqr, sequential/Nash: exactly max_outer Jacobi rounds (BASELINE configs[3])."""
    x0 = np.random.default_rng(8).uniform(-1, 1, (6, 16, 4))
    ref = oracle.strategy_run_batch(oracle.STRATEGY_SEQUENTIAL, oracle.MODEL_LQR, x0, max_outer=10, max_iterations=100, tolerance=1e-5,
                                    trig=oracle.TRIG_PORTABLE)
    got = mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, mas.example_desc(2), mas.IlqrParams.make(100, 1e-5), 10, x0)
    for k in ("trace_iters", "trace_cost", "costs", "total_cost", "X", "U"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("model,agents", [(1, 3), (1, 7), (2, 4)])
def test_line_search_nash(mas, ctx, oracle, model, agents):
    """LineSearchNashStrategy (strategies/nash.hpp:92-180): Jacobi solve, then a joint backtracking step
    old + alpha (cand - old) whenever the scenario's summed cost did not drop."""
    S = 12
    rng = np.random.default_rng(30 + agents)
    if model == 1:
        x0 = circ_x0(S, agents)[0]
        x0[:, :, 2] += rng.uniform(-0.2, 0.2, (S, agents))
        x0[:, :, 3] += rng.uniform(-1.0, 1.0, (S, agents))
    else:
        x0 = rng.uniform(-1, 1, (S, agents, 4))
    ref = oracle.strategy_run_batch(oracle.STRATEGY_LINESEARCH, model, x0, max_outer=6, max_iterations=100, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = mas.strategy_run(ctx, mas.Strategy.LINESEARCH, mas.example_desc(model), mas.IlqrParams.make(100, 1e-5), 6, x0)
    for k in ("trace_iters", "trace_cost", "costs", "total_cost", "X", "U"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("model,agents", [(1, 3), (1, 10), (2, 6)])
def test_centralized(mas, ctx, oracle, model, agents):
    """CentralizedStrategy (BASELINE configs[4] shape, smaller): stacked all-FD solve, per-agent cost re-evaluation."""
    S = 5
    rng = np.random.default_rng(agents)
    if model == 1:
        x0 = circ_x0(S, agents, jitter_seed=None)[0]
        x0[:, :, 2] += rng.uniform(-0.05, 0.05, (S, agents))
    else:
        x0 = rng.uniform(-1, 1, (S, agents, 4))
    ref = oracle.strategy_run_batch(oracle.STRATEGY_CENTRALIZED, model, x0, max_outer=1, max_iterations=100, tolerance=1e-5,
                                    trig=oracle.TRIG_PORTABLE)
    got = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(model), mas.IlqrParams.make(100, 1e-5), 1, x0)
    assert np.array_equal(got["trace_iters"][:, 0, 0], ref["trace_iters"][:, 0, 0])
    for k in ("total_cost", "costs", "X", "U"):
        assert np.array_equal(got[k], ref[k]), k


def test_unsupported_strategies_fail_loudly(mas, ctx):
    x0, _ = circ_x0(1, 3)
    for kind in (7,):
        with pytest.raises(mas.MasB200Error) as e:
            mas.strategy_run(ctx, kind, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 2, x0)
        assert e.value.code == 1


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_constrained_agents_inside_strategies(mas, ctx, oracle, kind):
    """Agents with path constraints (model 5) through sequential / line-search / trust-region Nash: the device batch keeps
    every agent's multipliers and penalty from round to round like the reference's per-agent solver clones
    (nash.hpp:17-21,76-84); bit-exact against the oracle, which tests/test_ref_pin.py pins to the reference's own code."""
    from conftest import random_x0

    x0 = random_x0(5, 24, seed=2).reshape(8, 3, 4)
    desc = mas.example_desc(5)
    strategy = {1: mas.Strategy.SEQUENTIAL, 2: mas.Strategy.LINESEARCH, 3: mas.Strategy.TRUSTREGION}[kind]
    got = mas.strategy_run(ctx, strategy, desc, mas.IlqrParams.make(6, 1e-5), 3, x0)
    prm = np.tile(np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5]), (8, 3, 1))
    ref = oracle.strategy_run_batch(kind, 5, x0, params=prm, max_outer=3, max_iterations=6, trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["trace_iters"], ref["trace_iters"])
    # a second run starts from fresh solvers again
    again = mas.strategy_run(ctx, strategy, desc, mas.IlqrParams.make(6, 1e-5), 3, x0)
    assert np.array_equal(again["U"], got["U"])


def test_centralized_time_budget(mas, ctx):
    """max_ms on the stacked solve (ilqr.hpp:84-90): a budget that is already spent stops before the first backward pass."""
    x0, _ = circ_x0(2, 4)
    prm = mas.IlqrParams.make(100, 1e-5, max_ms=-1.0)
    got = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(1), prm, 1, x0)
    assert (got["trace_iters"][:, 0, 0] == 0).all()
    assert (got["U"] == 0.0).all()
    free = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5, max_ms=1e9), 1, x0)
    ref = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 1, x0)
    assert np.array_equal(free["U"], ref["U"]) and (free["trace_iters"][:, 0, 0] > 0).all()
    # max_outer = 0 with trace pointers: nothing is written (the trace has no elements)
    z = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 0, x0)
    assert np.array_equal(z["U"], ref["U"]) and z["trace_iters"].size == 0


MIXED_MODELS = [0, 2, 3, 1, 4]  # ST-lane (4x2, T 80), LQR (4x4, T 10), pendulum (2x1, T 60), ST-circ (4x2, T 10), rocket (3x1, T 50)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_mixed_agents_nash(mas, ctx, oracle, kind):
    """mas_b200_strategy_run_mixed: agents of different registered models, shapes and horizons in one scenario; every Nash
    strategy, bit-exact against the oracle (pinned to the reference's own code for the same mix in tests/test_ref_pin.py)."""
    from conftest import random_x0

    S = 5
    x0 = [random_x0(m, S, seed=10 + m) for m in MIXED_MODELS]
    descs = [mas.example_desc(m) for m in MIXED_MODELS]
    U0 = [np.broadcast_to(mas.example_controls(m, d.horizon_steps), (S, d.horizon_steps, d.control_dim)).copy() for m, d in zip(MIXED_MODELS, descs)]
    strategy = {1: mas.Strategy.SEQUENTIAL, 2: mas.Strategy.LINESEARCH, 3: mas.Strategy.TRUSTREGION}[kind]
    got = mas.strategy_run_mixed(ctx, strategy, descs, mas.IlqrParams.make(8, 1e-5), 3, x0, U_init=U0)
    ref = oracle.strategy_run_mixed(kind, MIXED_MODELS, x0, max_outer=3, max_iterations=8, trig=oracle.TRIG_PORTABLE)
    for a in range(len(MIXED_MODELS)):
        assert np.array_equal(got["X"][a], ref["X"][a]), a
        assert np.array_equal(got["U"][a], ref["U"][a]), a
    assert np.array_equal(got["costs"], ref["costs"]) and np.array_equal(got["total_cost"], ref["total_cost"])
    assert np.array_equal(got["trace_iters"].sum(1), ref["iterations_total"])
    # all agents of one description: the same entry point takes the single-batch path
    d1 = [mas.example_desc(1)] * 3
    x1, _ = circ_x0(4, 3)
    same = mas.strategy_run_mixed(ctx, strategy, d1, mas.IlqrParams.make(100, 1e-5), 4, [x1[:, a] for a in range(3)])
    flat = mas.strategy_run(ctx, strategy, mas.example_desc(1), mas.IlqrParams.make(100, 1e-5), 4, x1)
    assert np.array_equal(same["costs"], flat["costs"]) and all(np.array_equal(same["U"][a], flat["U"][:, a]) for a in range(3))


@pytest.mark.parametrize("models", [[3, 4], [1, 0, 3, 2, 4], [2, 1, 2], [4, 3, 1]])
def test_mixed_agents_centralized(mas, ctx, oracle, models):
    """CentralizedStrategy over agents of different models and shapes (stacked_mixed.cuh; one CTA per scenario, run-time block
    structure, every derivative of the stacked functions by finite differences): stacked trajectories, per-agent costs, total
    and iteration count bit-equal to the oracle's build_global_ocp + iLQR, which tests/test_ref_pin.py pins to the reference's
    own code for the same mixes.  Result horizon = the first agent's (multi_agent_problem.hpp:65-69)."""
    from conftest import random_x0

    S = 3
    x0 = [random_x0(m, S, seed=40 + m) for m in models]
    descs = [mas.example_desc(m) for m in models]
    got = mas.strategy_run_mixed(ctx, mas.Strategy.CENTRALIZED, descs, mas.IlqrParams.make(6, 1e-5), 1, x0)
    ref = oracle.strategy_run_mixed(0, models, x0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    T0 = descs[0].horizon_steps
    for a in range(len(models)):
        assert got["X"][a].shape == (S, T0 + 1, descs[a].state_dim)
        assert np.array_equal(got["X"][a], ref["X"][a]), a
        assert np.array_equal(got["U"][a], ref["U"][a]), a
    assert np.array_equal(got["costs"], ref["costs"]) and np.array_equal(got["total_cost"], ref["total_cost"])
    assert np.array_equal(got["trace_iters"][:, 0, 0], ref["iterations_total"][:, 0])


def test_centralized_stack_above_256_states(mas, ctx, oracle):
    """mas_b200_strategy_run(CENTRALIZED) with 65 LQR agents (260 stacked states and controls): stacks too large for the
    compiled-in kernel of centralized.cuh go to the general solve in HBM (stacked_mixed.cuh) instead of
    MAS_B200_ERR_UNSUPPORTED.  Horizon 2, one iteration (the oracle needs 0.8 M stacked cost calls per step for the Hessians)."""
    from conftest import random_x0

    S, A, T = 2, 65, 2
    x0 = random_x0(2, S * A, seed=77).reshape(S, A, 4)
    desc = mas.example_desc(2)
    desc.horizon_steps = T
    got = mas.strategy_run(ctx, mas.Strategy.CENTRALIZED, desc, mas.IlqrParams.make(1, 1e-5), 1, x0)
    ref = oracle.strategy_run_batch(0, 2, x0, horizon=T, max_outer=1, max_iterations=1, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    for k in ("X", "U", "costs", "total_cost"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["trace_iters"][:, 0, 0], ref["trace_iters"][:, 0, 0]) and np.abs(got["U"]).max() > 0


def test_mixed_agents_stacked_functions(mas, ctx, oracle):
    """mas_b200_global_ocp_eval_mixed on the device: compute_offsets + build_global_ocp of mixed agents, what the reference's
    tests/ocp_tests.cpp:76-154 checks (ids out of order, offsets, bounds only if all agents have them, stacked values)."""
    from conftest import random_x0

    for models in (MIXED_MODELS, [3, 4, 0]):
        descs = [mas.example_desc(m) for m in models]
        nx, nu = sum(d.state_dim for d in descs), sum(d.control_dim for d in descs)
        X, U = np.linspace(0.1, 1.5, nx), np.linspace(-0.3, 0.3, nu)
        x0 = [random_x0(m, 1, seed=20 + m)[0] for m in models]
        ref = oracle.global_ocp_eval_mixed(models, x0, X, U)
        # agents handed over in reverse order with their ids: the blocks come out id-sorted
        A = len(models)
        got = mas.global_ocp_eval_mixed(ctx, descs[::-1], agent_ids=list(range(A))[::-1], X=X, U=U, time_index=3)
        assert list(got["block_agent"]) == list(range(A))[::-1]
        for k in ("total_x", "total_u", "horizon", "has_bounds", "dt", "stage", "terminal"):
            assert got[k] == ref[k], k
        assert np.array_equal(got["dynamics"], ref["dynamics"])
        if ref["has_bounds"]:
            assert np.array_equal(got["bounds"], ref["bounds"])
        assert list(got["state_offsets"]) == list(np.cumsum([0] + [d.state_dim for d in descs[:-1]]))
