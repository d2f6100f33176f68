"""PIN: the restated oracle (oracle/ref_*.hpp) against a build of the reference's OWN sources.

oracle/_ref/libref.so = /root/reference/include/multi_agent_solver/** + examples/*.cpp, unmodified,
compiled against oracle/eigen_shim (the image has no Eigen; the shim's arithmetic order is documented in
its header and is the one oracle/dense.hpp states).  Everything here is asserted BIT FOR BIT, in the
reference's own libm mode and in the portable-trig mode the GPU kernels are compared in (the reference's
sin / cos / tan calls are bound at link time, nothing in its source is touched): trajectories, costs,
and the counters the oracle defines (iterations, status, line-search candidates, Q_uu regularisation
retries), which the wrapper derives from the OCP's public callbacks.

libref.so can be built only where /root/reference exists; elsewhere the prebuilt file (it travels with
the snapshot) is used, and the tests skip if neither is there.
"""
import numpy as np
import pytest

from conftest import circle_x0, random_x0


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py

    if not ref_py.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference sources to build it from")
    ref_py.build()
    return ref_py


SOLVE_KEYS = ("X", "U", "cost", "iterations", "status", "rollouts", "alpha_trials", "reg_retries")


def assert_same(a, b, keys):
    for k in keys:
        assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), k


def test_reference_own_unit_tests_pass_on_the_shim(ref):
    """tests/ocp_tests.cpp of the reference, unmodified (a gtest stand-in provides TEST / EXPECT_*)."""
    if not ref.can_build():
        pytest.skip("needs /root/reference")
    out = ref.run_reference_unit_tests()
    assert "4 tests, 0 failed" in out


@pytest.mark.parametrize("trig", [0, 1])
def test_config1_single_track_ocp(oracle, ref, trig):
    """single_track_ocp --solver ilqr: 3 iterations, 18 candidates, 508.5930603049 (SURVEY 9 P2 confirmed
    by the reference's own code)."""
    x0 = np.array([[0.0, 1.0, 0.0, 0.0]])
    a = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, trig=trig)
    b = ref.ilqr_solve_batch(ref.MODEL_ST_LANE, x0, trig=trig)
    assert_same(a, b, SOLVE_KEYS)
    assert b["iterations"][0] == 3 and b["alpha_trials"][0] == 18 and b["status"][0] == 0
    assert abs(b["cost"][0] - 508.5930603049) < 1e-9


@pytest.mark.parametrize("trig", [0, 1])
def test_config3_batch_sample(oracle, ref, trig):
    """1,024 of the headline batch's problems (same generator ranges)."""
    x0 = random_x0(0, 1024, seed=11)
    a = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, x0, trig=trig)
    b = ref.ilqr_solve_batch(ref.MODEL_ST_LANE, x0, trig=trig)
    assert_same(a, b, SOLVE_KEYS)
    assert len(set(b["iterations"])) > 2


@pytest.mark.parametrize("model,max_it,tol", [(1, 100, 1e-5), (2, 100, 1e-5), (3, 1000, 1e-4), (4, 25, 1e-6)])
@pytest.mark.parametrize("trig", [0, 1])
def test_other_example_models(oracle, ref, model, max_it, tol, trig):
    """ST-circ (all FD, regularisation retries), LQR, pendulum (time-varying cost, m = 1), rocket."""
    x0 = random_x0(model, 6, seed=5 + model)
    U = np.broadcast_to(ref.default_controls(model), (6,) + ref.default_controls(model).shape).copy()
    assert np.array_equal(ref.default_controls(model), oracle.default_controls(model))
    a = oracle.ilqr_solve_batch(model, x0, U_init=U, max_iterations=max_it, tolerance=tol, trig=trig)
    b = ref.ilqr_solve_batch(model, x0, U_init=U, max_iterations=max_it, tolerance=tol, trig=trig)
    assert_same(a, b, SOLVE_KEYS)
    if model == 1:
        assert b["reg_retries"].max() > 0


def test_iteration_cap_and_warm_start(oracle, ref):
    x0 = random_x0(0, 8, seed=3)
    a = oracle.ilqr_solve_batch(0, x0, max_iterations=2, trig=1)
    b = ref.ilqr_solve_batch(0, x0, max_iterations=2, trig=1)
    assert_same(a, b, SOLVE_KEYS)
    assert (b["status"] == 1).any()  # MAX_ITER
    a2 = oracle.ilqr_solve_batch(0, x0, U_init=a["U"], max_iterations=10, trig=1)
    b2 = ref.ilqr_solve_batch(0, x0, U_init=b["U"], max_iterations=10, trig=1)
    assert_same(a2, b2, SOLVE_KEYS)


@pytest.mark.parametrize("trig", [0, 1])
def test_augmented_lagrangian_branch(oracle, ref, trig):
    """ilqr.hpp:121-170,236-260,380-407 with path constraints given through the OCP's public members; the same
    solver object solving four times (multipliers and penalty persist, :331-338)."""
    prm = np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5])
    x0 = np.array([0.0, 1.0, 0.0, 0.5])
    a = oracle.ilqr_solve_repeat(5, x0, 4, params=prm, trig=trig, max_iterations=10)
    b = ref.ilqr_solve_repeat(5, x0, 4, params=prm, trig=trig, max_iterations=10)
    assert_same(a, b, ("X", "U", "cost", "iterations"))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
@pytest.mark.parametrize("trig", [0, 1])
def test_strategies_circular_track(oracle, ref, kind, trig):
    """config 2 (trust region, 3 agents) and the other three strategies on the same problem; two scenarios,
    the second with a jittered track radius."""
    x0 = np.stack([circle_x0(3, 20.0), circle_x0(3, 17.5)])
    params = np.zeros((2, 3, 2))
    params[0] = [20.0, 5.0]
    params[1] = [17.5, 5.0]
    a = oracle.strategy_run_batch(kind, oracle.MODEL_ST_CIRC, x0, params=params, trig=trig)
    b = ref.strategy_run_batch(kind, ref.MODEL_ST_CIRC, x0, params=params, trig=trig)
    assert_same(a, b, ("X", "U", "costs", "total_cost"))
    its = a["trace_iters"].sum(1) if kind else a["trace_iters"][:, 0, :]
    assert np.array_equal(its, b["iterations_total"])


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_strategies_lqr(oracle, ref, kind):
    """config 4 (sequential, LQR agents) and the other strategies."""
    x0 = np.tile([1.0, 0, 0, 0], (1, 8, 1))
    a = oracle.strategy_run_batch(kind, oracle.MODEL_LQR, x0)
    b = ref.strategy_run_batch(kind, ref.MODEL_LQR, x0)
    assert_same(a, b, ("X", "U", "costs", "total_cost"))
    its = a["trace_iters"].sum(1) if kind else a["trace_iters"][:, 0, :]
    assert np.array_equal(its, b["iterations_total"])


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_strategies_with_constrained_agents(oracle, ref, kind):
    """Path constraints inside the Nash strategies: every agent keeps its own solver for all outer rounds
    (nash.hpp:17-21,76-84), so multipliers and penalty persist from round to round."""
    x0 = random_x0(5, 6, seed=2).reshape(2, 3, 4)
    prm = np.tile(np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5]), (2, 3, 1))
    a = oracle.strategy_run_batch(kind, 5, x0, params=prm, max_outer=3, max_iterations=6, trig=1)
    b = ref.strategy_run_batch(kind, 5, x0, params=prm, max_outer=3, max_iterations=6, trig=1)
    assert_same(a, b, ("X", "U", "costs", "total_cost"))
    assert np.array_equal(a["trace_iters"].sum(1), b["iterations_total"])


def test_config5_centralized_32_agents(oracle, ref):
    """build_global_ocp (multi_agent_problem.hpp:52-127) + centralized.hpp:18-38 + the stacked all-FD solve at
    n = 128, m = 64 (about 20 s of CPU for the two runs)."""
    x0 = circle_x0(32)[None]
    a = oracle.strategy_run_batch(0, oracle.MODEL_ST_CIRC, x0, trig=1)
    b = ref.strategy_run_batch(0, ref.MODEL_ST_CIRC, x0, trig=1, count_iterations=False)
    assert_same(a, b, ("X", "U", "costs", "total_cost"))


MIXED_MODELS = [0, 2, 3, 1, 4]  # ST-lane (4x2), LQR (4x4), pendulum (2x1), ST-circ (4x2), rocket (3x1)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_mixed_agents_nash(oracle, ref, kind):
    """Agents of different models, dims and horizons in one MultiAgentProblem (the reference takes any mix:
    multi_agent_problem.hpp:37-50; its own test stacks a 2x1 and a 1x2 agent)."""
    x0 = [random_x0(m, 3, seed=10 + m) for m in MIXED_MODELS]
    a = oracle.strategy_run_mixed(kind, MIXED_MODELS, x0, max_outer=3, max_iterations=8, trig=1)
    b = ref.strategy_run_mixed(kind, MIXED_MODELS, x0, max_outer=3, max_iterations=8, trig=1)
    for k in ("X", "U"):
        for xa, xb in zip(a[k], b[k]):
            assert np.array_equal(xa, xb), k
    assert_same(a, b, ("costs", "total_cost", "iterations_total"))


@pytest.mark.parametrize("models", [[3, 4], [1, 0, 3, 2, 4], [2, 1, 2], [4, 3, 1]])
def test_mixed_agents_centralized(oracle, ref, models):
    """CentralizedStrategy (centralized.hpp:18-38) on agents of different models: build_global_ocp of the mix, iLQR on the
    stacked OCP with every derivative by finite differences, every agent's rows and own objective back."""
    x0 = [random_x0(m, 2, seed=40 + m) for m in models]
    a = oracle.strategy_run_mixed(0, models, x0, max_iterations=6, trig=1)
    b = ref.strategy_run_mixed(0, models, x0, max_iterations=6, trig=1)
    for k in ("X", "U"):
        for xa, xb in zip(a[k], b[k]):
            assert np.array_equal(xa, xb), k
    assert_same(a, b, ("costs", "total_cost", "iterations_total"))


def test_mixed_agents_build_global_ocp(oracle, ref):
    """compute_offsets + build_global_ocp on mixed agents (tests/ocp_tests.cpp:76-154 does this for a 2x1 and a 1x2 agent):
    dims, horizon / dt of the first block, bounds only when every agent has both, block-diagonal dynamics, block-order sums."""
    for models in (MIXED_MODELS, [3, 4, 0]):
        x0 = [random_x0(m, 1, seed=20 + m)[0] for m in models]
        nx = sum(oracle.model_dims(m)[0] for m in models)
        nu = sum(oracle.model_dims(m)[1] for m in models)
        X, U = np.linspace(0.1, 1.5, nx), np.linspace(-0.3, 0.3, nu)
        a = oracle.global_ocp_eval_mixed(models, x0, X, U)
        b = ref.global_ocp_eval_mixed(models, x0, X, U)
        for k in ("total_x", "total_u", "horizon", "has_bounds", "dt", "stage", "terminal"):
            assert a[k] == b[k], k
        assert np.array_equal(a["dynamics"], b["dynamics"])
        assert np.array_equal(a["bounds"], b["bounds"], equal_nan=True)
    assert not a["has_bounds"] or True


@pytest.mark.parametrize("jm", [15, 6])
def test_analytic_constraint_jacobians(oracle, ref, jm):
    """OCP::*_constraints_*_jacobian given by the problem (ocp.hpp:65-68, used at ilqr.hpp:124-133,146-153)."""
    prm = np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5, float(jm)])
    x0 = np.array([0.0, 1.0, 0.0, 0.5])
    a = oracle.ilqr_solve_repeat(5, x0, 3, params=prm, trig=1, max_iterations=10)
    b = ref.ilqr_solve_repeat(5, x0, 3, params=prm, trig=1, max_iterations=10)
    assert_same(a, b, ("X", "U", "cost", "iterations"))
