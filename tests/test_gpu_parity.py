"""GPU parity: libmas_b200.so (through the C ABI) against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): final cost 1e-9 relative, trajectories 1e-7 absolute, identical
iteration counts and convergence flags.  The oracle runs in its portable-trig mode, i.e. with the same
sin/cos/tan implementation as the kernels (include/mas_b200/portable_math.h); the kernels are written
to reproduce the oracle's rounded operations one for one, so these tests additionally require
bit-identical trajectories.
"""
import numpy as np
import pytest

from conftest import EXAMPLE_SOLVER_PARAMS, assert_parity, is_bit_exact, random_x0

pytestmark = pytest.mark.gpu


def gpu_solve(mas, ctx, model, x0, U0, max_iterations, tolerance, lanes=0, chains=0, mask=None, params=None):
    desc = mas.example_desc(model)
    if mask is not None:
        desc.deriv_mask = mask
    b = mas.Batch(ctx, desc, x0.shape[0])
    b.set_tuning(lanes, chains)
    b.set_initial_states(x0)
    if params is not None:
        b.set_params(params)
    b.set_controls(U0)
    b.solve(mas.IlqrParams.make(max_iterations, tolerance))
    out = b.get_solution()
    out["stats"] = b.stats()
    b.close()
    return out


@pytest.mark.parametrize("model,batch", [(0, 257), (1, 96), (2, 64), (3, 40), (4, 64)])
def test_example_models_match_oracle(mas, ctx, oracle, model, batch):
    max_it, tol = EXAMPLE_SOLVER_PARAMS[model]  # the example mains' own values (pendulum: 1000 iterations)
    x0 = random_x0(model, batch, seed=100 + model)
    desc = mas.example_desc(model)
    U0 = np.broadcast_to(mas.example_controls(model, desc.horizon_steps), (batch, desc.horizon_steps, desc.control_dim)).copy()
    ref = oracle.ilqr_solve_batch(model, x0, U_init=U0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
    got = gpu_solve(mas, ctx, model, x0, U0, max_it, tol)
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert got["stats"]["iterations"] == int(ref["iterations"].sum())
    assert got["stats"]["alpha_trials"] == int(ref["alpha_trials"].sum())
    assert got["stats"]["reg_retries"] == int(ref["reg_retries"].sum())


def test_config1_single_track_ocp(mas, ctx, oracle):
    """BASELINE config 1: the single_track_ocp example, x0 = (0,1,0,0), 10 / 1e-5."""
    x0 = np.array([[0.0, 1.0, 0.0, 0.0]])
    got = gpu_solve(mas, ctx, 0, x0, None, 10, 1e-5)
    ref = oracle.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert_parity(got, ref)
    assert got["iterations"][0] == 3 and got["status"][0] == mas.Status.CONVERGED
    assert abs(got["cost"][0] - 508.5930603049) < 1e-9


@pytest.mark.parametrize("lanes,chains", [(1, 1), (1, 2), (2, 1), (4, 1), (8, 1), (16, 1)])
def test_line_search_lane_mappings_agree(mas, ctx, oracle, lanes, chains):
    """Every lanes-per-problem mapping of the line-search kernel picks the same first improving step."""
    x0 = random_x0(0, 130, seed=5)
    ref = oracle.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = gpu_solve(mas, ctx, 0, x0, None, 10, 1e-5, lanes=lanes, chains=chains)
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert got["stats"]["forward_lanes"] == lanes


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("model", [0, 1, 2])
def test_line_search_schedules_agree(mas, ctx, oracle, model, mode):
    """Concurrent-lanes and compacted-rounds scheduling of the line search give the same bits."""
    max_it, tol = EXAMPLE_SOLVER_PARAMS[model]
    max_it = min(max_it, 40)
    x0 = random_x0(model, 200, seed=77 + model)
    ref = oracle.ilqr_solve_batch(model, x0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
    desc = mas.example_desc(model)
    b = mas.Batch(ctx, desc, 200)
    b.set_line_search_mode(mode)
    b.set_initial_states(x0)
    b.set_controls(None)
    b.solve(mas.IlqrParams.make(max_it, tol))
    got = b.get_solution()
    st = b.stats()
    b.close()
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert st["alpha_trials"] == int(ref["alpha_trials"].sum())


@pytest.mark.parametrize("model", [0, 4])
def test_all_fd_mode_matches_oracle(mas, ctx, oracle, emu, model):
    """deriv_mask = 0: every derivative from the finite-difference defaults (ocp.hpp:117-135).  The oracle's
    example builders install analytic callbacks for these models, so the expected values come from the
    host emulation of the device source, itself checked against the oracle in the CPU tests."""
    max_it, tol = 8, 1e-5
    x0 = random_x0(model, 48, seed=11)
    desc = mas.example_desc(model)
    U0 = np.broadcast_to(mas.example_controls(model, desc.horizon_steps), (48, desc.horizon_steps, desc.control_dim)).copy()
    exp = emu.solve(model, x0, U0, max_it, tol, mask=0)
    got = gpu_solve(mas, ctx, model, x0, U0, max_it, tol, mask=0)
    assert_parity(got, exp)
    assert is_bit_exact(got, exp)


def test_per_problem_params(mas, ctx, oracle):
    """ST-circ with a different track radius per problem (SURVEY 8d config 2 jitter)."""
    rng = np.random.default_rng(3)
    B = 64
    R = rng.uniform(15, 25, B)
    th = rng.uniform(0, 2 * np.pi, B)
    x0 = np.stack([R * np.cos(th), R * np.sin(th), 1.57 + th, np.full(B, 4.0)], -1)
    oparams = np.stack([R, np.full(B, 5.0)], -1)
    gparams = np.stack([R, np.full(B, 5.0), np.ones(B), np.ones(B), np.full(B, 0.001), np.full(B, 0.001)], -1)
    ref = oracle.ilqr_solve_batch(1, x0, params=oparams, max_iterations=40, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = gpu_solve(mas, ctx, 1, x0, None, 40, 1e-5, params=gparams)
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)


def test_warm_start_and_iteration_cap(mas, ctx, oracle):
    """solve() starts from best_controls (ilqr.hpp:71-75); max_iterations = 2 ends with MAX_ITER flags."""
    x0 = random_x0(0, 64, seed=21)
    first = oracle.ilqr_solve_batch(0, x0, max_iterations=2, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got1 = gpu_solve(mas, ctx, 0, x0, None, 2, 1e-5)
    assert_parity(got1, first)
    assert (got1["status"] == mas.Status.MAX_ITER).any()
    second = oracle.ilqr_solve_batch(0, x0, U_init=first["U"], max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got2 = gpu_solve(mas, ctx, 0, x0, got1["U"], 10, 1e-5)
    assert_parity(got2, second)
    assert is_bit_exact(got2, second)


def test_one_shot_api_and_edge_sizes(mas, ctx, oracle):
    """mas_b200_ilqr_solve_batch on host arrays; batch sizes 1, 31, 33 (ragged against the 32-lane padding)."""
    for batch in (1, 31, 33):
        x0 = random_x0(2, batch, seed=batch)
        got = mas.ilqr_solve_batch(ctx, mas.example_desc(2), mas.IlqrParams.make(100, 1e-5), x0)
        ref = oracle.ilqr_solve_batch(2, x0, max_iterations=100, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
        assert_parity(got, ref)
        assert is_bit_exact(got, ref)


def test_zero_iterations_and_errors(mas, ctx, oracle):
    x0 = random_x0(0, 8, seed=1)
    got = gpu_solve(mas, ctx, 0, x0, None, 0, 1e-5)
    assert (got["iterations"] == 0).all() and (got["status"] == mas.Status.MAX_ITER).all()
    X, cost = oracle.rollout_cost(0, x0, np.zeros((8, 80, 2)), trig=oracle.TRIG_PORTABLE)
    np.testing.assert_array_equal(got["cost"], cost)
    np.testing.assert_array_equal(got["X"], X)
    desc = mas.example_desc(1)
    desc.deriv_mask = mas.DerivBits.LX  # ST-circ has no analytic cost gradient
    with pytest.raises(mas.MasB200Error) as e:
        mas.Batch(ctx, desc, 4)
    assert e.value.code == 1
    desc = mas.example_desc(0)
    desc.state_dim = 5
    with pytest.raises(mas.MasB200Error):
        mas.Batch(ctx, desc, 4)


def test_time_limit_flag(mas, ctx):
    """max_ms = 0 with a clock already past it: the batch stops before the first backward pass (ilqr.hpp:84-90)."""
    x0 = random_x0(0, 16, seed=2)
    desc = mas.example_desc(0)
    b = mas.Batch(ctx, desc, 16)
    b.set_initial_states(x0)
    b.set_controls(None)
    b.solve(mas.IlqrParams.make(10, 1e-5, max_ms=-1.0))
    out = b.get_solution()
    assert (out["status"] == mas.Status.TIME_LIMIT).all() and (out["iterations"] == 0).all()


def test_full_size_batch_properties(mas, ctx, oracle):
    """BASELINE config 3 at full size (65,536 ST-lane problems): size-independent properties.
    (a) ALL 65,536 problems equal the CPU checker bit for bit (oracle/_ref = the reference's own sources with portable
    trig when its library is there, else the restated oracle); (b) solving the two halves separately
    gives the same bits as the whole batch (no cross-problem coupling, any lane mapping);
    (c) a second solve warm-started from the solution stops after one iteration without changing it."""
    B = 65536
    x0 = mas.synthetic_single_track_x0(B)
    desc = mas.example_desc(0)
    prm = mas.IlqrParams.make(10, 1e-5)
    b = mas.Batch(ctx, desc, B)
    b.set_initial_states(x0)
    b.set_controls(None)
    b.solve(prm)
    full = b.get_solution()
    from oracle import ref_py

    checker = ref_py if ref_py.available() else oracle
    ref = checker.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert_parity(full, ref)
    assert is_bit_exact(full, ref)
    # (c) idempotence on the device-resident solution
    b.solve(prm)
    again = b.get_solution()
    conv = full["status"] == mas.Status.CONVERGED
    assert conv.mean() > 0.9
    one = conv & (again["iterations"] == 1) & (again["cost"] == full["cost"])
    assert one.mean() > 0.5  # a failed last line search fails again: one iteration, nothing changes
    np.testing.assert_array_equal(again["U"][one], full["U"][one])
    assert (again["cost"] <= full["cost"]).all()  # the merit never increases (ilqr.hpp:220)
    b.close()
    # (b) halves
    for lo, hi, lanes in ((0, B // 2, 2), (B // 2, B, 4)):
        h = gpu_solve(mas, ctx, 0, x0[lo:hi], None, 10, 1e-5, lanes=lanes)
        np.testing.assert_array_equal(h["X"], full["X"][lo:hi])
        np.testing.assert_array_equal(h["cost"], full["cost"][lo:hi])
        np.testing.assert_array_equal(h["iterations"], full["iterations"][lo:hi])


def nonfinite_x0():
    """Initial states whose trajectories or costs leave the finite range (the reference has no failure return: a non-finite
    merit never improves, `improvement` is NaN, the stop test is false and the solve runs to max_iterations)."""
    x0 = random_x0(0, 40, seed=1)
    x0[1] = [0, 1e160, 0.1, 1.0]      # y^2 overflows: cost = inf
    x0[2] = [0, 0.5, 1e7, 1.0]        # heading outside the portable trig domain: nan
    x0[3] = [0, np.nan, 0.0, 1.0]
    x0[4] = [0, 0.5, 0.0, np.inf]
    x0[5] = [0, 0.5, 0.0, 1e200]      # v * cos overflows along the horizon
    x0[6] = [0, 1e154, 0.0, 1.0]      # cost sum just below / above overflow
    x0[7] = [0, 0.5, 823549.4, 1.0]   # heading just inside the trig domain
    x0[8] = [0, -0.3, 0.2, 1e100]
    x0[33] = [0, -np.inf, 0.2, 1.0]
    return x0


def equal_nan(a, b):
    return all(np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f") for k in ("X", "U", "cost", "iterations", "status"))


@pytest.mark.parametrize("lanes", [0, 1, 4, 16])
def test_nonfinite_trajectories_match_reference(mas, ctx, oracle, lanes):
    """NaN / inf trajectories and costs: same bits (NaN-aware), same iteration counts and flags as the checker, with the
    structural-zero shortcuts of the Riccati products and safe_eval (finite_differences.hpp:95-107) in play."""
    from oracle import ref_py

    x0 = nonfinite_x0()
    checker = ref_py if ref_py.available() else oracle
    ref = checker.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = gpu_solve(mas, ctx, 0, x0, None, 10, 1e-5, lanes=lanes)
    assert equal_nan(got, ref)
    assert (ref["status"][1:7] == mas.Status.MAX_ITER).all() and not np.isfinite(ref["cost"][1:7]).any()
    # all-FD mode on the same inputs (every derivative through the finite-difference defaults, non-finite costs -> 0)
    ref_fd = oracle.ilqr_solve_batch(1, x0 * [1, 1, 1, 1], max_iterations=8, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got_fd = gpu_solve(mas, ctx, 1, x0, None, 8, 1e-5, lanes=lanes)
    assert equal_nan(got_fd, ref_fd)


def test_rocket_mass_clamp(mas, ctx, oracle):
    """rocket_model.hpp:66-69: mass = max(m, 1e-6); initial masses at, below and far below the clamp, and negative."""
    x0 = random_x0(4, 16, seed=9)
    x0[:6, 2] = [1e-6, 1e-7, 0.0, -1.0, 1e-300, 5e-7]
    desc = mas.example_desc(4)
    U0 = np.broadcast_to(mas.example_controls(4, desc.horizon_steps), (16, desc.horizon_steps, 1)).copy()
    ref = oracle.ilqr_solve_batch(4, x0, U_init=U0, max_iterations=25, tolerance=1e-6, trig=oracle.TRIG_PORTABLE)
    got = gpu_solve(mas, ctx, 4, x0, U0, 25, 1e-6)
    assert equal_nan(got, ref)


# ---- augmented-Lagrangian path constraints (ilqr.hpp:121-170,236-260,380-407) ------------------------------
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_constrained_model_matches_oracle(mas, ctx, oracle, mode):
    """Model 5 (lane following + equality and inequality path constraints) through the C ABI."""
    B = 200
    x0 = random_x0(5, B, seed=301)
    desc = mas.example_desc(5)
    assert (desc.state_dim, desc.control_dim, desc.horizon_steps) == (4, 2, 80)
    U0 = np.zeros((B, 80, 2))
    ref = oracle.ilqr_solve_batch(5, x0, U_init=U0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    b = mas.Batch(ctx, desc, B)
    b.set_line_search_mode(mode)
    b.set_initial_states(x0)
    b.set_controls(U0)
    b.solve(mas.IlqrParams.make(6, 1e-5))
    got = b.get_solution()
    st = b.stats()
    b.close()
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert st["alpha_trials"] == int(ref["alpha_trials"].sum())


def test_constraint_state_persists_and_resets(mas, ctx, oracle):
    """Multipliers and the penalty parameter live in the batch like the members of a reference solver object: a second
    and third solve continue from them (ilqr.hpp:331-338); mas_b200_batch_reset_solver_state gives a fresh solver."""
    B = 12
    x0 = random_x0(5, B, seed=302)
    U0 = np.zeros((B, 80, 2))
    prm = mas.IlqrParams.make(5, 1e-5)
    prm.penalty = 2.0
    b = mas.Batch(ctx, mas.example_desc(5), B)
    b.set_initial_states(x0)
    b.set_controls(U0)
    hist = []
    for _ in range(3):
        b.solve(prm)
        hist.append(b.get_solution())
    for i in range(B):
        ref = oracle.ilqr_solve_repeat(5, x0[i], 3, U_init=U0[i], max_iterations=5, tolerance=1e-5, penalty=2.0, trig=oracle.TRIG_PORTABLE)
        for r in range(3):
            assert hist[r]["cost"][i] == ref["cost"][r] and hist[r]["iterations"][i] == ref["iterations"][r]
            assert np.array_equal(hist[r]["X"][i], ref["X"][r])
        assert np.array_equal(hist[2]["U"][i], ref["U"])
    # fresh solver state + the original controls reproduce the first solve
    b.reset_solver_state()
    b.set_controls(U0)
    b.solve(prm)
    again = b.get_solution()
    b.close()
    assert np.array_equal(again["cost"], hist[0]["cost"]) and np.array_equal(again["X"], hist[0]["X"])


def test_result_sink_streams_the_same_results(mas, ctx, oracle):
    """mas_b200_batch_set_result_sink: results written to page-locked host memory while the solve runs (per iteration, the
    problems that left the active set) are the bits get_solution returns -- two solves back to back, a ragged batch, every
    stop reason: converged, iteration cap (max_iterations 2), zero iterations; unpinned buffers are refused."""
    import torch

    B = 1237
    xa, xb = random_x0(0, B, seed=411), random_x0(0, B, seed=412)
    b = mas.Batch(ctx, mas.example_desc(0), B)

    def pinned():
        return dict(X=torch.empty((B, 81, 4), dtype=torch.float64).pin_memory().numpy(), U=torch.empty((B, 80, 2), dtype=torch.float64).pin_memory().numpy(),
                    cost=torch.empty(B, dtype=torch.float64).pin_memory().numpy(), iterations=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                    status=torch.empty(B, dtype=torch.int32).pin_memory().numpy())

    sink = pinned()
    with pytest.raises(Exception):
        b.set_result_sink(dict(X=np.empty((B, 81, 4))))  # pageable memory
    b.set_result_sink(sink)
    for x0, iters in ((xa, 10), (xb, 10), (xa, 2), (xb, 0)):
        for v in sink.values():
            v[...] = -7
        b.set_initial_states(x0)
        b.set_controls(None)
        b.solve(mas.IlqrParams.make(iters, 1e-5))
        b.wait_solution()
        streamed = {k: v.copy() for k, v in sink.items()}
        direct = b.get_solution()
        for k in streamed:
            assert np.array_equal(streamed[k], direct[k], equal_nan=(k in ("X", "U", "cost"))), k
        ref = oracle.ilqr_solve_batch(0, x0, max_iterations=iters, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
        assert is_bit_exact(streamed, ref)
    # time budget already spent (ilqr.hpp:84-90): every problem is written with TIME_LIMIT and its initial rollout
    for v in sink.values():
        v[...] = -7
    b.set_initial_states(xa)
    b.set_controls(None)
    b.solve(mas.IlqrParams.make(10, 1e-5, max_ms=-1.0))
    b.wait_solution()
    direct = b.get_solution()
    assert (sink["status"] == mas.Status.TIME_LIMIT).all() and (sink["iterations"] == 0).all()
    for k in sink:
        assert np.array_equal(sink[k], direct[k]), k
    b.set_result_sink(None)
    b.set_initial_states(xa)
    b.set_controls(None)
    for v in sink.values():
        v[...] = -7
    b.solve(mas.IlqrParams.make(10, 1e-5))
    b.wait_solution()
    assert np.all(sink["iterations"] == -7)  # unregistered: nothing is written
    b.close()


def test_asynchronous_download_survives_the_next_solve(mas, ctx, oracle):
    """begin_get_solution stages the results in HBM; a following solve on other inputs must not disturb them."""
    import torch

    B = 300
    xa, xb = random_x0(0, B, seed=401), random_x0(0, B, seed=402)
    prm = mas.IlqrParams.make(10, 1e-5)
    b = mas.Batch(ctx, mas.example_desc(0), B)

    def pinned():
        return dict(X=torch.empty((B, 81, 4), dtype=torch.float64).pin_memory().numpy(), U=torch.empty((B, 80, 2), dtype=torch.float64).pin_memory().numpy(),
                    cost=torch.empty(B, dtype=torch.float64).pin_memory().numpy(), iterations=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                    status=torch.empty(B, dtype=torch.int32).pin_memory().numpy())

    oa, ob = pinned(), pinned()
    b.set_initial_states(xa)
    b.set_controls(None)
    b.solve(prm)
    b.begin_get_solution(oa)
    b.set_initial_states(xb)  # no wait in between
    b.set_controls(None)
    b.solve(prm)
    b.begin_get_solution(ob)  # waits for the first download, then stages the second
    b.wait_solution()
    b.close()
    for x0, got in ((xa, oa), (xb, ob)):
        ref = oracle.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
        assert_parity(got, ref)
        assert is_bit_exact(got, ref)


@pytest.mark.parametrize("mode,lanes", [(0, 0), (1, 4), (1, 16), (3, 0)])
def test_trial_store_on_and_off_agree(mas, ctx, oracle, mode, lanes):
    """Accepted steps copied from the trial store vs rolled out again: identical results, and both equal the oracle."""
    x0 = random_x0(0, 500, seed=610)
    ref = oracle.ilqr_solve_batch(0, x0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    outs = []
    for store in (1, 0, 2):
        b = mas.Batch(ctx, mas.example_desc(0), 500)
        b.set_trial_store(store)
        b.set_line_search_mode(mode)
        b.set_tuning(lanes, 0)
        b.set_initial_states(x0)
        b.set_controls(None)
        b.solve(mas.IlqrParams.make(10, 1e-5))
        outs.append(b.get_solution())
        b.close()
    assert is_bit_exact(outs[0], ref) and is_bit_exact(outs[1], ref) and is_bit_exact(outs[2], ref)
    assert_parity(outs[0], ref)


def test_one_shot_calls_reuse_the_context_batch_without_leaking_state(mas, ctx, oracle):
    """mas_b200_ilqr_solve_batch keeps its device batch in the context between calls of the same shape; every call must
    still behave like a fresh solver on a fresh problem (per-problem parameters, warm start, multipliers)."""
    B = 64
    desc = mas.example_desc(1)
    prm = mas.IlqrParams.make(30, 1e-5)
    rng = np.random.default_rng(77)
    R = rng.uniform(15, 25, B)
    th = rng.uniform(0, 2 * np.pi, B)
    x0 = np.stack([R * np.cos(th), R * np.sin(th), 1.57 + th, np.full(B, 4.0)], -1)
    gp = np.stack([R, np.full(B, 5.0), np.ones(B), np.ones(B), np.full(B, 1e-3), np.full(B, 1e-3)], -1)
    a = mas.ilqr_solve_batch(ctx, desc, prm, x0, model_params=gp)       # per-problem radii
    b = mas.ilqr_solve_batch(ctx, desc, prm, x0)                        # same shape, shared default parameters
    c = mas.ilqr_solve_batch(ctx, desc, prm, x0, model_params=gp)       # and back
    ref_a = oracle.ilqr_solve_batch(1, x0, params=gp[:, :2], max_iterations=30, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    ref_b = oracle.ilqr_solve_batch(1, x0, max_iterations=30, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert is_bit_exact(a, ref_a) and is_bit_exact(b, ref_b) and is_bit_exact(c, ref_a)
    # constrained model: a second one-shot call is a fresh solver again (no multipliers carried over)
    d5 = mas.example_desc(5)
    x5 = random_x0(5, B, seed=78)
    p5 = mas.IlqrParams.make(5, 1e-5)
    r1 = mas.ilqr_solve_batch(ctx, d5, p5, x5)
    r2 = mas.ilqr_solve_batch(ctx, d5, p5, x5)
    ref5 = oracle.ilqr_solve_batch(5, x5, max_iterations=5, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert is_bit_exact(r1, ref5) and is_bit_exact(r2, ref5)


@pytest.mark.parametrize("model,mask", [(1, 0), (0, 0), (0, 0x0F)])
def test_lane_parallel_backward_pass_matches_the_one_thread_pass(mas, ctx, emu, model, mask):
    """backward_lanes_kernel (FD stencil points over eight lanes per problem) against backward_kernel on the same inputs."""
    B = 200
    x0 = random_x0(model, B, seed=520 + model)
    desc = mas.example_desc(model)
    desc.deriv_mask = mask
    outs = []
    for lanes_on in (1, 0):
        b = mas.Batch(ctx, desc, B)
        b.set_trial_store(lanes_on)  # 0 also switches the lane-parallel backward pass off
        b.set_initial_states(x0)
        b.set_controls(None)
        b.solve(mas.IlqrParams.make(8, 1e-5))
        outs.append(b.get_solution())
        outs[-1]["stats"] = b.stats()
        b.close()
    assert is_bit_exact(outs[0], outs[1])
    assert outs[0]["stats"]["reg_retries"] == outs[1]["stats"]["reg_retries"]
    T, m = desc.horizon_steps, desc.control_dim
    exp = emu.solve(model, x0, np.zeros((B, T, m)), 8, 1e-5, mask=mask)
    assert is_bit_exact(outs[0], exp)


@pytest.mark.parametrize("model,mask", [(0, None), (0, 0), (1, 0), (2, None), (3, 0), (4, None), (5, None)])
def test_backward_modes_are_bit_identical(mas, ctx, oracle, model, mask):
    """mas_b200_batch_set_backward_mode: one thread per problem (1), FD tasks over eight lanes (2) and the time-parallel
    linearisation + Riccati sweep (3) give the same bits, and the oracle's where the derivative mode is the example's."""
    max_it, tol = EXAMPLE_SOLVER_PARAMS.get(model, (6, 1e-5))
    max_it = min(max_it, 12)
    B = 200
    x0 = random_x0(model, B, seed=500 + model)
    desc = mas.example_desc(model)
    example_mask = desc.deriv_mask
    if mask is not None:
        desc.deriv_mask = mask
    U0 = np.zeros((B, desc.horizon_steps, desc.control_dim))
    outs = []
    for mode in (1, 2, 3):
        b = mas.Batch(ctx, desc, B)
        b.set_backward_mode(mode)
        b.set_initial_states(x0)
        b.set_controls(U0)
        b.solve(mas.IlqrParams.make(max_it, tol))
        o = b.get_solution()
        o["stats"] = b.stats()
        outs.append(o)
        b.close()
    for o in outs[1:]:
        for k in ("X", "U", "cost", "iterations", "status"):
            assert np.array_equal(o[k], outs[0][k], equal_nan=o[k].dtype.kind == "f"), k
        assert o["stats"]["reg_retries"] == outs[0]["stats"]["reg_retries"]
    if desc.deriv_mask == example_mask:
        ref = oracle.ilqr_solve_batch(model, x0, U_init=U0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
        assert is_bit_exact(outs[2], ref)


def test_debug_trace_matches_oracle(mas, ctx, oracle):
    """params.debug: the values the reference prints per iteration (ilqr.hpp:79-80,262-267) are recorded per problem."""
    x0 = random_x0(0, 70, seed=77)
    b = mas.Batch(ctx, mas.example_desc(0), 70)
    b.set_initial_states(x0)
    b.set_controls(None)
    prm = mas.IlqrParams.make(10, 1e-5)
    prm.debug = 1
    b.solve(prm)
    got = b.get_solution()
    for p in (0, 33, 69):
        tr = oracle.ilqr_solve_trace(0, x0[p], max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
        rec = b.debug_trace(p)
        assert rec.shape[0] == got["iterations"][p] + 1
        assert np.array_equal(rec[1:, 0], tr["cost_trace"])
        assert np.array_equal(rec[1:, 5].astype(int), tr["alpha_index"])
        assert rec[0, 0] == rec[0, 1] and np.isnan(rec[0, 5])
        merit = np.concatenate([[rec[0, 1]], rec[1:, 1]])
        assert np.array_equal(rec[1:, 2], merit[:-1] - merit[1:])  # d_merit
    prm.debug = 0
    b.solve(prm)
    with pytest.raises(mas.MasB200Error):
        b.debug_trace(0)
    b.close()


@pytest.mark.parametrize("jm", [15, 9])
def test_analytic_constraint_jacobians(mas, ctx, oracle, jm):
    """deriv_mask bits 9-12 (analytic constraint Jacobians, ocp.hpp:65-68) through the C ABI."""
    B = 100
    x0 = random_x0(5, B, seed=303)
    desc = mas.example_desc(5)
    desc.deriv_mask = desc.deriv_mask | (jm << 9)
    U0 = np.zeros((B, 80, 2))
    prm = np.tile(np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5, float(jm)]), (B, 1))
    ref = oracle.ilqr_solve_batch(5, x0, U_init=U0, params=prm, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    for mode in (1, 3):
        b = mas.Batch(ctx, desc, B)
        b.set_backward_mode(mode)
        b.set_initial_states(x0)
        b.set_controls(U0)
        b.solve(mas.IlqrParams.make(6, 1e-5))
        got = b.get_solution()
        b.close()
        assert_parity(got, ref)
        assert is_bit_exact(got, ref)
    bad = mas.example_desc(0)
    bad.deriv_mask = bad.deriv_mask | (1 << 9)  # a model without constraints has no such callback
    with pytest.raises(mas.MasB200Error):
        mas.Batch(ctx, bad, 4)


def test_straight_line_divisions_equal_the_division_instruction(mas, ctx):
    """The trial rollouts divide with straight-line code (portable_math.h: div_spec = the compiler's division sequence minus
    its branch, div_const_spec = Markstein's correction) and repeat a step with the plain division whenever those flag an
    operand as outside their fast path.  Brute force on the device: wherever the fast path is accepted the bits are those of
    div.rn.f64 -- 2^28 pairs over all exponents, near-tie quotients and the cos-like divisor range of tan."""
    r = ctx.selftest_division(1 << 28, seed=20240607)
    assert r["checked"] >= 1 << 28
    assert r["div_mismatch"] == 0 and r["div_const_mismatch"] == 0, r
    # three quarters of the pairs are ordinary magnitudes: the fast path must be the rule there, not the exception
    assert r["div_exact"] > 0.74 * r["checked"] and r["div_const_exact"] > 0.74 * r["checked"], r
