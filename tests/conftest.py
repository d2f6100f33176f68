"""Shared fixtures.  GPU tests are marked `gpu`; everything else runs on a machine without one.

`/root/reference` is never read here: inputs are generated, expected values come from oracle/ (a CPU
restatement of the reference, test infrastructure) or from tests/golden/.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def mas():
    import multi_agent_solver_b200 as m

    m.load_library()
    return m


@pytest.fixture(scope="session")
def ctx(mas):
    return mas.Context(0)


# ---- host emulation of the device source (tests/csrc/host_emulation.cpp) -------------------------
_EMU_SRC = os.path.join(ROOT, "tests", "csrc", "host_emulation.cpp")
_EMU_LIB = os.path.join(ROOT, "tests", "_build", "libhost_emulation.so")


def _build_emulation() -> str:
    # MAS_B200_EMU_SANITIZE=1: the same source with -fsanitize=address,undefined into a second library (run the emulation
    # tests with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): index errors of the device
    # code that would be silent on the GPU; tools/README.md
    if os.environ.get("MAS_B200_EMU_SANITIZE") == "1":
        return _build_emulation_variant(_EMU_LIB.replace(".so", "_asan.so"), ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])
    return _build_emulation_variant(_EMU_LIB, ["-O2"])


def _build_emulation_variant(_EMU_LIB: str, opt) -> str:
    deps = [_EMU_SRC] + [os.path.join(ROOT, "multi_agent_solver_b200", "csrc", f) for f in ("ilqr_core.cuh", "models.cuh", "centralized.cuh", "stacked_mixed.cuh")]
    deps.append(os.path.join(ROOT, "include", "mas_b200", "portable_math.h"))
    if not os.path.exists(_EMU_LIB) or os.path.getmtime(_EMU_LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(_EMU_LIB), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", *opt, "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-x", "c++",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "multi_agent_solver_b200", "csrc"), _EMU_SRC,
                               "-o", _EMU_LIB])
    return _EMU_LIB


# model id -> (n, m, T, dt, example mask, has_bounds, lower, upper, default params)
MODEL_TABLE = {
    0: (4, 2, 80, 0.1, 0x3F, 1, [-0.7, -1.0], [0.7, 1.0], [1.0, 10.0, 1.0, 0.1, 0.1]),
    1: (4, 2, 10, 0.5, 0x0, 1, [-0.5, -0.5], [0.5, 0.5], [20.0, 5.0, 1.0, 1.0, 0.001, 0.001]),
    2: (4, 4, 10, 0.1, 0x1FF, 0, [0.0] * 4, [0.0] * 4, []),
    3: (2, 1, 60, 0.05, 0x0, 1, [-5.0], [5.0], [60.0]),
    4: (3, 1, 50, 0.1, 0x1BF, 1, [0.0], [20.0], [9.81, 50.0, 5e-3, 15.0, 2.0, 0.0]),
    # lane following + path constraints (equality a = k (v_des - v), inequality v <= v_max): not a reference example,
    # exercises the augmented-Lagrangian branch of iLQR::solve
    5: (4, 2, 80, 0.1, 0x3F, 1, [-0.7, -1.0], [0.7, 1.0], [1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5]),
}


class HostEmulation:
    def __init__(self):
        self.lib = ctypes.CDLL(_build_emulation())

    def solve(self, model, x0, U, max_iterations, tolerance, L=1, C=2, mask=None, per_problem_params=None, penalty=10.0, repeats=1,
              trial_store=True, backward_lanes=0, sweep_lanes=False, sweep_wide=0):
        """repeats > 1: the same solver state (multipliers, penalty) and warm start solving again; adds cost_history /
        iterations_history [repeats, B]."""
        n, m, T, dt, emask, hb, lo, hi, prm = MODEL_TABLE[model]
        mask = emask if mask is None else mask
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        B = x0.shape[0]
        U = np.array(U, dtype=np.float64, order="C").reshape(B, T, m).copy()
        X = np.zeros((B, T + 1, n))
        cost = np.zeros(B)
        it = np.zeros(B, np.int32)
        st = np.zeros(B, np.int32)
        tr = np.zeros(B, np.int32)
        rg = np.zeros(B, np.int32)
        lo = np.array(list(lo) + [0.0] * 8)
        hi = np.array(list(hi) + [0.0] * 8)
        sp = np.array(list(prm) + [0.0] * 8)
        P = ctypes.POINTER(ctypes.c_double)
        PI = ctypes.POINTER(ctypes.c_int)
        pp = None
        if per_problem_params is not None:
            ppa = np.ascontiguousarray(per_problem_params, dtype=np.float64)
            pp = ppa.ctypes.data_as(P)
        self.lib.emu_set_trial_store(int(trial_store))
        self.lib.emu_set_backward_lanes(int(backward_lanes))
        self.lib.emu_set_sweep_lanes(int(bool(sweep_lanes)))
        self.lib.emu_set_sweep_wide(int(sweep_wide))
        hc = np.zeros((repeats, B))
        hi_ = np.zeros((repeats, B), np.int32)
        self.lib.emu_set_al_options(ctypes.c_double(penalty), ctypes.c_double(5.0), ctypes.c_double(1e-4), ctypes.c_double(1e-6), int(repeats),
                                    hc.ctypes.data_as(P), hi_.ctypes.data_as(PI))
        rc = self.lib.emu_ilqr_solve_batch(model, B, T, ctypes.c_double(dt), ctypes.c_uint(mask), hb, lo.ctypes.data_as(P), hi.ctypes.data_as(P),
                                           sp.ctypes.data_as(P), pp, x0.ctypes.data_as(P), U.ctypes.data_as(P), X.ctypes.data_as(P),
                                           cost.ctypes.data_as(P), it.ctypes.data_as(PI), st.ctypes.data_as(PI), tr.ctypes.data_as(PI),
                                           rg.ctypes.data_as(PI), int(max_iterations), ctypes.c_double(tolerance), int(L), int(C))
        self.lib.emu_set_al_options(ctypes.c_double(10.0), ctypes.c_double(5.0), ctypes.c_double(1e-4), ctypes.c_double(1e-6), 1, None, None)
        self.lib.emu_set_backward_lanes(0)
        self.lib.emu_set_sweep_lanes(0)
        self.lib.emu_set_sweep_wide(0)
        assert rc == 0
        return dict(X=X, U=U, cost=cost, iterations=it, status=st, alpha_trials=tr, reg_retries=rg, cost_history=hc, iterations_history=hi_)


    def solve_centralized(self, model, x0, max_iterations=100, tolerance=1e-5):
        """stacked_solve<M> (centralized.cuh) with tid = 0, nthr = 1.  x0: [agents, n]."""
        n, m, T, dt, _, hb, lo, hi, prm = MODEL_TABLE[model]
        A = x0.shape[0]
        ns, ms = A * n, A * m
        U = np.zeros((T, ms))
        X = np.zeros((T + 1, ns))
        oc = np.zeros(1 + A)
        oi = np.zeros(4, np.int32)
        pp = np.tile(np.array(list(prm) + [0.0])[: max(len(prm), 1)], (A, 1)).copy()
        lo = np.array(list(lo) + [0.0] * 8)
        hi = np.array(list(hi) + [0.0] * 8)
        x0f = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1)
        P = ctypes.POINTER(ctypes.c_double)
        rc = self.lib.emu_centralized_solve(model, A, T, ctypes.c_double(dt), hb, lo.ctypes.data_as(P), hi.ctypes.data_as(P), pp.ctypes.data_as(P),
                                            x0f.ctypes.data_as(P), U.ctypes.data_as(P), X.ctypes.data_as(P), oc.ctypes.data_as(P),
                                            oi.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), int(max_iterations), ctypes.c_double(tolerance))
        assert rc == 0
        return dict(X=X.reshape(T + 1, A, n).transpose(1, 0, 2).copy(), U=U.reshape(T, A, m).transpose(1, 0, 2).copy(), total_cost=oc[0],
                    costs=oc[1:].copy(), iterations=int(oi[0]), status=int(oi[1]), reg_retries=int(oi[2]), alpha_trials=int(oi[3]))


    def solve_centralized_mixed(self, models, x0_list, max_iterations=100, tolerance=1e-5, horizon=0):
        """mixed_stacked_solve (stacked_mixed.cuh) with tid = 0, nthr = 1: one scenario, agents of different models.
        x0_list[a]: [n_a].  Horizon / dt of the first agent (horizon > 0 overrides the example's), bounds only if every agent
        has them (build_global_ocp)."""
        rows = [MODEL_TABLE[m] for m in models]
        A = len(models)
        T, dt = horizon or rows[0][2], rows[0][3]
        ns, ms = sum(r[0] for r in rows), sum(r[1] for r in rows)
        hb = int(all(r[5] for r in rows))
        lo = np.concatenate([np.asarray(r[6], dtype=np.float64) for r in rows])
        hi = np.concatenate([np.asarray(r[7], dtype=np.float64) for r in rows])
        pp = np.zeros((A, 8))
        for a, r in enumerate(rows):
            pp[a, :len(r[8])] = r[8]
        x0 = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in x0_list]))
        X = np.zeros((T + 1, ns))
        U = np.zeros((T, ms))
        oc = np.zeros(1 + A)
        oi = np.zeros(4, np.int32)
        marr = np.array(models, dtype=np.int32)
        P = ctypes.POINTER(ctypes.c_double)
        PI = ctypes.POINTER(ctypes.c_int)
        rc = self.lib.emu_centralized_mixed(A, marr.ctypes.data_as(PI), T, ctypes.c_double(dt), hb, lo.ctypes.data_as(P), hi.ctypes.data_as(P),
                                            pp.ctypes.data_as(P), x0.ctypes.data_as(P), X.ctypes.data_as(P), U.ctypes.data_as(P), oc.ctypes.data_as(P),
                                            oi.ctypes.data_as(PI), int(max_iterations), ctypes.c_double(tolerance))
        assert rc == 0
        Xs, Us, ox, ou = [], [], 0, 0
        for r in rows:
            Xs.append(X[:, ox:ox + r[0]].copy())
            Us.append(U[:, ou:ou + r[1]].copy())
            ox += r[0]
            ou += r[1]
        return dict(X=Xs, U=Us, total_cost=oc[0], costs=oc[1:].copy(), iterations=int(oi[0]), status=int(oi[1]), reg_retries=int(oi[2]),
                    alpha_trials=int(oi[3]))


@pytest.fixture(scope="session")
def emu():
    return HostEmulation()


def circle_x0(n_agents: int, radius: float = 20.0) -> np.ndarray:
    """Agents evenly spaced on the circular track (multi_agent_single_track.cpp:41-44,114-119)."""
    th = 2.0 * np.pi * np.arange(n_agents) / n_agents
    return np.stack([radius * np.cos(th), radius * np.sin(th), 1.57 + th, np.full(n_agents, 4.0)], -1)


# ---- synthetic inputs ------------------------------------------------------------------------------
def random_x0(model: int, batch: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if model in (0, 5):  # config 3 ranges
        return np.stack([np.zeros(batch), rng.uniform(-2, 2, batch), rng.uniform(-0.5, 0.5, batch), rng.uniform(0, 2, batch)], -1)
    if model == 1:  # agents on the circle, tangential heading (multi_agent_single_track.cpp:41-44)
        th = rng.uniform(0, 2 * np.pi, batch)
        return np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(batch, 4.0)], -1)
    if model == 2:
        return rng.uniform(-1, 1, (batch, 4))
    if model == 3:
        return np.stack([np.pi - 0.05 + rng.uniform(-0.1, 0.1, batch), rng.uniform(-0.1, 0.1, batch)], -1)
    if model == 4:
        return np.stack([rng.uniform(0, 1, batch), rng.uniform(-1, 1, batch), rng.uniform(0.8, 1.2, batch)], -1)
    raise ValueError(model)


# solver params of the example mains: (max_iterations, tolerance)
EXAMPLE_SOLVER_PARAMS = {0: (10, 1e-5), 1: (100, 1e-5), 2: (100, 1e-5), 3: (1000, 1e-4), 4: (25, 1e-6)}


def assert_parity(got, ref, cost_rtol=1e-9, traj_atol=1e-7):
    """BASELINE.json north_star tolerances: cost 1e-9 relative, trajectories 1e-7 absolute, identical
    iteration counts and convergence flags."""
    np.testing.assert_array_equal(got["iterations"], ref["iterations"])
    np.testing.assert_array_equal(got["status"], ref["status"])
    denom = np.maximum(np.abs(ref["cost"]), 1e-300)
    assert np.max(np.abs(got["cost"] - ref["cost"]) / denom) <= cost_rtol
    assert np.max(np.abs(got["X"] - ref["X"])) <= traj_atol
    assert np.max(np.abs(got["U"] - ref["U"])) <= traj_atol


def is_bit_exact(got, ref) -> bool:
    return bool(np.array_equal(got["X"], ref["X"]) and np.array_equal(got["U"], ref["U"]) and np.array_equal(got["cost"], ref["cost"]))
