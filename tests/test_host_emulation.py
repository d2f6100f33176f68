"""The device source (multi_agent_solver_b200/csrc/ilqr_core.cuh + models.cuh, all __host__ __device__)
compiled with g++ and driven with the engine's schedule, against the oracle in portable-trig mode.
Bit-identical results are required: the kernels restate the oracle's rounded operations one for one.
This is the CPU-side check of the CUDA code; the same comparison runs on the GPU in test_gpu_parity.py.
"""
import numpy as np
import pytest

from conftest import EXAMPLE_SOLVER_PARAMS, MODEL_TABLE, assert_parity, is_bit_exact, random_x0


def _defaults(oracle, model, batch):
    n, m, T = MODEL_TABLE[model][:3]
    return np.broadcast_to(oracle.default_controls(model), (batch, T, m)).copy()


@pytest.mark.parametrize("model,batch,cap", [(0, 48, 10), (1, 12, 100), (2, 16, 100), (3, 6, 40), (4, 12, 25)])
def test_models_bit_exact(emu, oracle, model, batch, cap):
    max_it, tol = EXAMPLE_SOLVER_PARAMS[model]
    max_it = min(max_it, cap)
    x0 = random_x0(model, batch, seed=40 + model)
    U0 = _defaults(oracle, model, batch)
    ref = oracle.ilqr_solve_batch(model, x0, U_init=U0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(model, x0, U0, max_it, tol)
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert np.array_equal(got["alpha_trials"], ref["alpha_trials"])
    assert np.array_equal(got["reg_retries"], ref["reg_retries"])


@pytest.mark.parametrize("L,C", [(1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (4, 2), (8, 1), (8, 2), (16, 1), (32, 1), (32, 2)])
def test_lane_mappings_pick_the_first_improving_step(emu, oracle, L, C):
    x0 = random_x0(0, 24, seed=9)
    U0 = np.zeros((24, 80, 2))
    ref = oracle.ilqr_solve_batch(0, x0, U_init=U0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(0, x0, U0, 10, 1e-5, L=L, C=C)
    assert is_bit_exact(got, ref)
    assert np.array_equal(got["iterations"], ref["iterations"]) and np.array_equal(got["alpha_trials"], ref["alpha_trials"])


def test_regularisation_path(emu, oracle):
    """ST-circ from rest-like states exercises Q_uu + reg*I retries (ilqr.hpp:175-182)."""
    x0 = random_x0(1, 8, seed=3)
    U0 = np.zeros((8, 10, 2))
    ref = oracle.ilqr_solve_batch(1, x0, U_init=U0, max_iterations=30, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(1, x0, U0, 30, 1e-5)
    assert ref["reg_retries"].sum() > 0
    assert np.array_equal(got["reg_retries"], ref["reg_retries"])
    assert is_bit_exact(got, ref)


def test_runtime_mask_equals_compile_time_mask(emu):
    """backward_thread<M, -1> (mask read at run time) and the specialised instantiation agree."""
    x0 = random_x0(0, 8, seed=12)
    U0 = np.zeros((8, 80, 2))
    a = emu.solve(0, x0, U0, 6, 1e-5, mask=0x3F)          # specialised (example mask)
    b = emu.solve(0, x0, U0, 6, 1e-5, mask=0x3F & ~0x20)  # generic path, l_uu from finite differences
    c = emu.solve(0, x0, U0, 6, 1e-5, mask=0)             # all finite differences
    assert np.array_equal(a["iterations"] > 0, np.ones(8, bool))
    # l_uu is constant for this cost, its FD estimate is close but not equal: results differ slightly
    assert np.max(np.abs(a["cost"] - b["cost"]) / a["cost"]) < 1e-2
    assert np.max(np.abs(a["cost"] - c["cost"]) / a["cost"]) < 1e-1


def test_per_problem_params(emu, oracle):
    rng = np.random.default_rng(5)
    B = 10
    R = rng.uniform(15, 25, B)
    th = rng.uniform(0, 2 * np.pi, B)
    x0 = np.stack([R * np.cos(th), R * np.sin(th), 1.57 + th, np.full(B, 4.0)], -1)
    ref = oracle.ilqr_solve_batch(1, x0, params=np.stack([R, np.full(B, 5.0)], -1), max_iterations=30, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    pp = np.stack([R, np.full(B, 5.0), np.ones(B), np.ones(B), np.full(B, 0.001), np.full(B, 0.001)], -1)
    got = emu.solve(1, x0, np.zeros((B, 10, 2)), 30, 1e-5, per_problem_params=pp)
    assert is_bit_exact(got, ref)


@pytest.mark.parametrize("model,agents", [(1, 3), (1, 8), (2, 5)])
def test_centralized_stacked_solve_bit_exact(emu, oracle, model, agents):
    """CentralizedStrategy: the structure-exploiting stacked solve (centralized.cuh) reproduces the oracle's
    dense evaluation of build_global_ocp + iLQR with all-FD derivatives bit for bit, noise included."""
    from conftest import circle_x0

    x0 = circle_x0(agents) if model == 1 else np.random.default_rng(0).uniform(-1, 1, (agents, 4))
    ref = oracle.strategy_run_batch(oracle.STRATEGY_CENTRALIZED, model, x0[None], max_outer=1, max_iterations=100, tolerance=1e-5,
                                    trig=oracle.TRIG_PORTABLE)
    got = emu.solve_centralized(model, x0)
    assert got["iterations"] == ref["trace_iters"][0, 0, 0]
    assert got["total_cost"] == ref["total_cost"][0]
    assert np.array_equal(got["costs"], ref["costs"][0])
    assert np.array_equal(got["X"], ref["X"][0]) and np.array_equal(got["U"], ref["U"][0])


def test_nan_initial_state_matches_oracle(emu, oracle):
    """Non-finite costs propagate silently (SURVEY 8b error conventions): a NaN merit never accepts a
    candidate and `improvement < tolerance` is false for NaN, so the loop runs to max_iterations."""
    x0 = np.array([[0.0, np.nan, 0.0, 1.0]])
    U0 = np.zeros((1, 80, 2))
    ref = oracle.ilqr_solve_batch(0, x0, U_init=U0, max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(0, x0, U0, 10, 1e-5)
    assert got["iterations"][0] == ref["iterations"][0]
    assert got["status"][0] == ref["status"][0]
    assert np.isnan(got["cost"][0]) and np.isnan(ref["cost"][0])


# ---- augmented-Lagrangian path constraints (ilqr.hpp:121-170,236-260,380-407) ------------------------------
def test_constrained_model_bit_exact(emu, oracle):
    """Model 5 = lane following + one equality and one inequality path constraint: constraint Jacobians by finite
    differences, multiplier / penalty updates and the merit line search against the oracle."""
    B = 24
    x0 = random_x0(5, B, seed=77)
    U0 = _defaults(oracle, 5, B)
    ref = oracle.ilqr_solve_batch(5, x0, U_init=U0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(5, x0, U0, 6, 1e-5)
    assert_parity(got, ref)
    assert is_bit_exact(got, ref)
    assert np.array_equal(got["alpha_trials"], ref["alpha_trials"])
    # the constraints matter: the unconstrained problem from the same inputs ends elsewhere
    free = oracle.ilqr_solve_batch(0, x0, U_init=U0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert np.mean(np.abs(free["cost"] - ref["cost"]) > 1e-6) > 0.5


@pytest.mark.parametrize("L", [1, 4, 32])
def test_constrained_model_lane_mappings(emu, oracle, L):
    x0 = random_x0(5, 40, seed=78)
    U0 = np.zeros((40, 80, 2))
    ref = oracle.ilqr_solve_batch(5, x0, U_init=U0, max_iterations=4, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(5, x0, U0, 4, 1e-5, L=L, C=1)
    assert is_bit_exact(got, ref)
    assert np.array_equal(got["alpha_trials"], ref["alpha_trials"])


@pytest.mark.parametrize("penalty", [10.0, 0.5])
def test_multipliers_and_penalty_persist_across_solves(emu, oracle, penalty):
    """The same solver object solving again starts from its multipliers and its grown penalty (ilqr.hpp:331-338)."""
    x0 = random_x0(5, 6, seed=79)
    U0 = np.zeros((6, 80, 2))
    got = emu.solve(5, x0, U0, 5, 1e-5, penalty=penalty, repeats=3)
    changed = 0
    for b in range(6):
        ref = oracle.ilqr_solve_repeat(5, x0[b], 3, U_init=U0[b], max_iterations=5, tolerance=1e-5, penalty=penalty, trig=oracle.TRIG_PORTABLE)
        assert np.array_equal(got["cost_history"][:, b], ref["cost"])
        assert np.array_equal(got["iterations_history"][:, b], ref["iterations"])
        assert np.array_equal(got["U"][b], ref["U"]) and np.array_equal(got["X"][b], ref["X"][-1])
        changed += int(ref["cost"][1] != ref["cost"][0])
    assert changed > 0  # later solves do move: the persistent state is really in play


@pytest.mark.parametrize("model,L,C", [(0, 4, 1), (0, 4, 2), (0, 8, 2), (0, 16, 1), (5, 4, 2), (5, 16, 1), (1, 8, 1), (4, 16, 1)])
def test_stored_trials_equal_recomputed_steps(emu, model, L, C):
    """The accepted step copied from the trial store is the rollout the commit pass would repeat: same bits."""
    x0 = random_x0(model, 40, seed=90 + model)
    T, m = MODEL_TABLE[model][2], MODEL_TABLE[model][1]
    U0 = np.zeros((40, T, m))
    a = emu.solve(model, x0, U0, 6, 1e-5, L=L, C=C, trial_store=True)
    b = emu.solve(model, x0, U0, 6, 1e-5, L=L, C=C, trial_store=False)
    for k in ("X", "U", "cost", "iterations", "status", "alpha_trials"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("model,mask,lanes", [(1, 0, 8), (3, 0, 8), (0, 0, 8), (0, 0x3F & ~0x30, 4), (4, 0, 3), (5, 0, 8)])
def test_lane_parallel_backward_pass_is_bit_identical(emu, oracle, model, mask, lanes):
    """backward_lanes: FD stencil points dealt out as tasks to the lanes of a problem, Riccati on lane 0.  Every task
    repeats the arithmetic of the whole-matrix FD routines, so the one-thread backward pass is reproduced bit for bit
    (and with it the oracle where the oracle has the same derivative mode)."""
    max_it, tol = EXAMPLE_SOLVER_PARAMS.get(model, (6, 1e-5))
    max_it = min(max_it, 12)
    x0 = random_x0(model, 10, seed=300 + model)
    n, m, T = MODEL_TABLE[model][:3]
    U0 = np.zeros((10, T, m))
    one = emu.solve(model, x0, U0, max_it, tol, mask=mask)
    par = emu.solve(model, x0, U0, max_it, tol, mask=mask, backward_lanes=lanes)
    for k in ("X", "U", "cost", "iterations", "status", "alpha_trials", "reg_retries"):
        assert np.array_equal(one[k], par[k]), k
    if mask == MODEL_TABLE[model][4]:  # the example's own derivative mode: the oracle applies
        ref = oracle.ilqr_solve_batch(model, x0, U_init=U0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
        assert is_bit_exact(par, ref)


def test_nonfinite_trajectories(emu, oracle):
    """NaN / inf trajectories and costs through the device source on the CPU: same bits (NaN-aware), iteration counts and
    flags as the oracle and -- when its library is there -- as the reference's own sources (oracle/_ref)."""
    from oracle import ref_py

    x0 = random_x0(0, 24, seed=1)
    x0[1] = [0, 1e160, 0.1, 1.0]
    x0[2] = [0, 0.5, 1e7, 1.0]
    x0[3] = [0, np.nan, 0.0, 1.0]
    x0[4] = [0, 0.5, 0.0, np.inf]
    x0[5] = [0, 0.5, 0.0, 1e200]
    x0[6] = [0, 1e154, 0.0, 1.0]
    x0[7] = [0, 0.5, 823549.4, 1.0]
    x0[8] = [0, -0.3, 0.2, 1e100]
    U = np.zeros((24, 80, 2))
    refs = [oracle.ilqr_solve_batch(0, x0, U_init=U, trig=oracle.TRIG_PORTABLE)]
    if ref_py.available():
        refs.append(ref_py.ilqr_solve_batch(0, x0, U_init=U, trig=1))
    for L, C, bl, sl, sw in ((1, 2, 0, False, 0), (4, 1, 0, False, 0), (16, 1, 0, False, 0), (16, 1, -2, False, 0), (16, 1, -2, True, 0), (16, 1, -2, False, 2)):
        got = emu.solve(0, x0, U, 10, 1e-5, L=L, C=C, backward_lanes=bl, sweep_lanes=sl, sweep_wide=sw)
        for ref in refs:
            for k in ("X", "U", "cost", "iterations", "status"):
                assert np.array_equal(got[k], ref[k], equal_nan=got[k].dtype.kind == "f"), (L, C, k)
    assert not np.isfinite(refs[0]["cost"][1:7]).any() and (refs[0]["status"][1:7] == 1).all()
    # all-FD circular-track model on the same initial states
    Uc = np.zeros((24, 10, 2))
    ref_fd = oracle.ilqr_solve_batch(1, x0, U_init=Uc, max_iterations=8, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got_fd = emu.solve(1, x0, Uc, 8, 1e-5)
    for k in ("X", "U", "cost", "iterations", "status"):
        assert np.array_equal(got_fd[k], ref_fd[k], equal_nan=got_fd[k].dtype.kind == "f"), k


@pytest.mark.parametrize("model,mask,threads", [(0, None, 2), (0, 0, 8), (1, 0, 8), (2, None, 1), (3, 0, 3), (4, None, 2), (5, None, 2), (0, 0x3F & ~0x30, 5)])
def test_time_parallel_backward_pass_is_bit_identical(emu, oracle, model, mask, threads):
    """linearize_point for all (problem, t) pairs first -- each by `threads` cooperating threads --, then the Riccati
    recursion alone (riccati_sweep_thread): the derivative evaluation leaves the sequential chain (ilqr.hpp:106-113 is
    independent across t).  Same bits as the fused one-thread pass, and as the oracle in the example's derivative mode."""
    max_it, tol = EXAMPLE_SOLVER_PARAMS.get(model, (6, 1e-5))
    max_it = min(max_it, 12)
    x0 = random_x0(model, 10, seed=400 + model)
    n, m, T = MODEL_TABLE[model][:3]
    mask = MODEL_TABLE[model][4] if mask is None else mask
    U0 = np.zeros((10, T, m))
    one = emu.solve(model, x0, U0, max_it, tol, mask=mask)
    par = emu.solve(model, x0, U0, max_it, tol, mask=mask, backward_lanes=-threads)
    # ... and with the recursion itself column-parallel over the lanes of a problem (RiccatiLanes)
    lan = emu.solve(model, x0, U0, max_it, tol, mask=mask, backward_lanes=-threads, sweep_lanes=True)
    # ... and with one lane per entry of the NX x NX matrices (RiccatiWide; models it does not cover fall back), the lanes of
    # a phase run in ascending and in mixed order
    wides = [emu.solve(model, x0, U0, max_it, tol, mask=mask, backward_lanes=-threads, sweep_wide=w) for w in (1, 2)]
    for k in ("X", "U", "cost", "iterations", "status", "alpha_trials", "reg_retries"):
        assert np.array_equal(one[k], par[k]), k
        assert np.array_equal(one[k], lan[k]), k
        for wd in wides:
            assert np.array_equal(one[k], wd[k]), k
    if mask == MODEL_TABLE[model][4]:
        ref = oracle.ilqr_solve_batch(model, x0, U_init=U0, max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_PORTABLE)
        assert is_bit_exact(par, ref)
        assert is_bit_exact(lan, ref)


@pytest.mark.parametrize("jm", [15, 5, 10])
def test_analytic_constraint_jacobians(emu, oracle, jm):
    """deriv_mask bits 9-12: the constraint Jacobians a problem installs itself (ocp.hpp:65-68) instead of the
    finite-difference defaults (ocp.hpp:137-171), any subset; bit-exact against the oracle given the same callbacks."""
    x0 = random_x0(5, 6, seed=41)
    U0 = np.zeros((6, 80, 2))
    prm = np.tile(np.array([1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5, float(jm)]), (6, 1))
    ref = oracle.ilqr_solve_batch(5, x0, U_init=U0, params=prm, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    fd = oracle.ilqr_solve_batch(5, x0, U_init=U0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve(5, x0, U0, 6, 1e-5, mask=0x3F | (jm << 9))
    assert is_bit_exact(got, ref) and np.array_equal(got["iterations"], ref["iterations"])
    assert not np.array_equal(ref["U"], fd["U"])  # the analytic Jacobians do change the rounding


@pytest.mark.parametrize("models", [[3, 4], [1, 0, 3, 2, 4], [2, 1, 2], [4, 3, 1]])
def test_centralized_over_mixed_agents_is_bit_identical(emu, oracle, models):
    """CentralizedStrategy on agents of different models and shapes (stacked_mixed.cuh, the device source on the host) against the
    oracle's stacked solve -- which tests/test_ref_pin.py pins to the reference's own build_global_ocp + iLQR for the same mixes.
    [3, 4] is the shape of the reference's own stacking test (a 2x1 and a 3x1 agent; tests/ocp_tests.cpp:76-154 uses 2x1 and 1x2)."""
    from conftest import random_x0

    x0 = [random_x0(m, 1, seed=40 + m) for m in models]
    ref = oracle.strategy_run_mixed(0, models, x0, max_iterations=6, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve_centralized_mixed(models, [x[0] for x in x0], max_iterations=6, tolerance=1e-5)
    for a in range(len(models)):
        assert np.array_equal(got["X"][a], ref["X"][a][0]), a
        assert np.array_equal(got["U"][a], ref["U"][a][0]), a
    assert np.array_equal(got["costs"], ref["costs"][0]) and got["total_cost"] == ref["total_cost"][0]
    assert got["iterations"] == ref["iterations_total"][0, 0]


def test_centralized_stack_above_256_states(emu, oracle):
    """65 LQR agents = 260 stacked states and controls: past the size the compiled-in kernel of centralized.cuh is laid out for,
    so the C ABI sends the stack to the general solve (stacked_mixed.cuh); the same source on the host against the oracle's
    CentralizedStrategy.  Horizon 2 and one iteration: a finite-difference Hessian of this size is 0.8 M stacked cost calls per step."""
    from conftest import random_x0

    A, T = 65, 2
    x0 = random_x0(2, A, seed=77)
    ref = oracle.strategy_run_batch(0, 2, x0[None], horizon=T, max_outer=1, max_iterations=1, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    got = emu.solve_centralized_mixed([2] * A, list(x0), max_iterations=1, tolerance=1e-5, horizon=T)
    assert np.array_equal(np.stack(got["X"]), ref["X"][0]) and np.array_equal(np.stack(got["U"]), ref["U"][0])
    assert np.array_equal(got["costs"], ref["costs"][0]) and got["total_cost"] == ref["total_cost"][0]
    assert got["iterations"] == ref["trace_iters"][0, 0, 0] == 1 and np.abs(got["U"][0]).max() > 0


def _race_check(source, exe_name, args, drop_var, drop_ks):
    """Builds tests/csrc/<source> with -fsanitize=thread, runs it clean (no report, results identical to one thread), then with
    single barriers dropped: the sanitizer / the comparison has to catch those."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_dir = os.path.join(root, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, exe_name)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-mfma", "-fsanitize=thread", "-w",
                           "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "multi_agent_solver_b200", "csrc"), "-x", "c++",
                           os.path.join(root, "tests", "csrc", source), "-o", exe, "-lpthread"])
    env = dict(os.environ)
    env.pop(drop_var, None)
    env.pop("LD_PRELOAD", None)
    clean = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900, env=env)
    if "FATAL: ThreadSanitizer" in clean.stderr:  # e.g. an address-space layout the runtime cannot map
        pytest.skip("ThreadSanitizer cannot run here: " + clean.stderr.strip().splitlines()[0])
    assert clean.returncode == 0 and "ALL OK" in clean.stdout and "ThreadSanitizer" not in clean.stderr, clean.stdout + clean.stderr
    caught = 0
    for k in drop_ks:
        env[drop_var] = str(k)
        mutant = subprocess.run([exe, *args], capture_output=True, text=True, timeout=900, env=env)
        caught += int(mutant.returncode != 0 and ("WARNING: ThreadSanitizer: data race" in mutant.stderr or "DIFFERS" in mutant.stdout))
    assert caught >= len(drop_ks) - 1, caught
    return clean.stdout


def test_mixed_stacked_solve_has_no_races_between_barriers():
    """stacked_mixed.cuh with real threads as the threads of a CTA (tests/csrc/mixed_threads_test.cpp): MAS_CTA_SYNC() becomes a
    pthread barrier, 2 / 5 / 8 threads run the same source under ThreadSanitizer on four agent mixes.  No report = every pair of
    conflicting accesses to the workspace and the result arrays is ordered by a barrier (what __syncthreads() must do on the
    GPU), and the results equal the one-thread run bit for bit.  The harness is checked on itself: with one barrier dropped
    (MIXED_DROP_BARRIER=k) the sanitizer has to report the race (152 of the first 160 barriers are caught when dropped; the
    rest are redundant ones, e.g. two barriers in a row)."""
    _race_check("mixed_threads_test.cpp", "mixed_threads_tsan", [], "MIXED_DROP_BARRIER", (40, 75, 100, 130))


def test_centralized_kernel_has_no_races_between_barriers():
    """centralized.cuh (config 5's kernel) the same way (tests/csrc/centralized_threads_test.cpp): 7 / 32 / 128 / 160 host threads;
    from 128 threads on the Q_uu factorisation (threads 0-63, named barrier 1) runs beside the next step's finite differences
    (the other threads, named barrier 2) exactly as in the CUDA kernel, the named barriers mapped to pthread barriers whose
    participant counts are checked.  4 stacked circular-track agents, 10 steps, 4 iterations with 18 regularisation retries.
    Dropped CTA barriers are caught 26 times out of 30 sampled."""
    out = _race_check("centralized_threads_test.cpp", "centralized_threads_tsan", ["4", "10", "4"], "CENTRALIZED_DROP_BARRIER", (41, 65, 89, 113))
    assert "retries 18" in out and out.count("identical to one thread") == 4


def test_lane_cooperative_backward_passes_have_no_races():
    """RiccatiLanes (the column-parallel Riccati sweep of small active sets: four lanes per problem exchanging columns through
    shared memory, a __syncwarp() between phases a..e) and backward_lanes (eight lanes dealing out the finite-difference
    stencil points) with one host thread per lane and pthread barriers placed as in riccati_sweep_lanes_kernel /
    backward_lanes_kernel, under ThreadSanitizer (tests/csrc/lanes_threads_test.cpp): no report, gains bit-equal to the
    one-thread sweep.  Every one of the first 60 barriers is caught when dropped."""
    out = _race_check("lanes_threads_test.cpp", "lanes_threads_tsan", [], "LANES_DROP_BARRIER", (2, 9, 23, 41))
    assert out.count("identical to the one-thread sweep") == 6
