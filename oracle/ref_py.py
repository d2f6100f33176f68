"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes front-end to oracle/_ref/libref.so.

libref.so is a build of the reference's OWN sources (/root/reference/include/multi_agent_solver/**,
examples/*.cpp; unmodified, read in place) against oracle/eigen_shim -- see oracle/ref/Makefile.  It
can only be BUILT where /root/reference exists (this container); the built file travels to the GPU
box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
_TESTS_PATH = os.path.join(_HERE, "_ref", "ref_ocp_tests")
REFERENCE_ROOT = os.environ.get("MAS_REFERENCE_ROOT", "/root/reference")

MODEL_ST_LANE, MODEL_ST_CIRC, MODEL_LQR, MODEL_PENDULUM, MODEL_ROCKET, MODEL_ST_LANE_CON = range(6)
STRATEGY_CENTRALIZED, STRATEGY_SEQUENTIAL, STRATEGY_LINESEARCH, STRATEGY_TRUSTREGION = range(4)
TRIG_GLIBC, TRIG_PORTABLE = 0, 1

_lib = None


def can_build() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "include", "multi_agent_solver"))


def available() -> bool:
    return os.path.exists(_LIB_PATH) or can_build()


def build(force: bool = False) -> str:
    """Compile the reference's sources with oracle/ref/Makefile (needs /root/reference; ~30 s)."""
    if can_build():
        cmd = ["make", "-C", os.path.join(_HERE, "ref"), "-j8", f"REF={REFERENCE_ROOT}", "all"] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    elif not os.path.exists(_LIB_PATH):
        raise RuntimeError("oracle/_ref/libref.so is absent and the reference sources are not here to build it")
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)  # RTLD_LOCAL: its sin/cos/tan definitions stay private to it
    return _lib


def run_reference_unit_tests() -> str:
    """The reference's tests/ocp_tests.cpp, unmodified, against the shim.  Returns its stdout; raises on failure."""
    build()
    return subprocess.run([_TESTS_PATH], check=True, capture_output=True, text=True).stdout


def _p(a, ct=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def model_dims(model: int, horizon: int = 0):
    n, m, T, dt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    if lib().ref_model_dims(model, horizon, ctypes.byref(n), ctypes.byref(m), ctypes.byref(T), ctypes.byref(dt)):
        raise ValueError(f"ref: unknown model {model}")
    return n.value, m.value, T.value, dt.value


def max_threads() -> int:
    return lib().ref_max_threads()


def synthetic_single_track_x0(batch: int, seed: int = 20240607) -> np.ndarray:
    """Config-3 initial states: std::mt19937_64(seed); Y, psi, v per problem."""
    x0 = np.empty((batch, 4))
    lib().ref_synthetic_single_track_x0(ctypes.c_ulonglong(seed), int(batch), _p(x0))
    return x0


def default_controls(model: int, horizon: int = 0) -> np.ndarray:
    n, m, T, _ = model_dims(model, horizon)
    U = np.zeros((T, m))
    if lib().ref_default_controls(model, horizon, _p(U)):
        raise RuntimeError("ref_default_controls failed")
    return U


def ilqr_solve_batch(model, x0, U_init=None, params=None, horizon=0, max_iterations=10, tolerance=1e-5, max_ms=float("inf"),
                     trig=TRIG_GLIBC, threads=0, instrument=True):
    """mas::solve(Solver&, OCP&) per problem.  Same dict as oracle_py.ilqr_solve_batch."""
    x0 = _f64(x0)
    batch = x0.shape[0]
    n, m, T, _ = model_dims(model, horizon)
    if U_init is None:
        U = np.broadcast_to(default_controls(model, horizon), (batch, T, m)).copy()
    else:
        U = np.array(U_init, dtype=np.float64, order="C").reshape(batch, T, m).copy()
    params = _f64(params)
    np_ = 0 if params is None else params.shape[1]
    X = np.zeros((batch, T + 1, n))
    cost = np.zeros(batch)
    iters = np.full(batch, -1, dtype=np.int32)
    status = np.full(batch, -1, dtype=np.int32)
    stats = np.full((batch, 3), -1, dtype=np.int32)
    rc = lib().ref_ilqr_solve_batch(
        model, batch, _p(x0), _p(params), np_, horizon, _p(U), int(max_iterations), ctypes.c_double(tolerance), ctypes.c_double(max_ms),
        int(trig), int(threads), int(bool(instrument)), _p(X), _p(cost), _p(iters, ctypes.c_int), _p(status, ctypes.c_int),
        _p(stats, ctypes.c_int))
    if rc:
        raise RuntimeError("ref_ilqr_solve_batch failed")
    return dict(X=X, U=U, cost=cost, iterations=iters, status=status, rollouts=stats[:, 0], alpha_trials=stats[:, 1], reg_retries=stats[:, 2])


def ilqr_solve_repeat(model, x0, n_repeat, U_init=None, params=None, horizon=0, max_iterations=10, tolerance=1e-5, penalty=10.0,
                      trig=TRIG_GLIBC):
    x0 = _f64(x0).reshape(-1)
    n, m, T, _ = model_dims(model, horizon)
    U = default_controls(model, horizon) if U_init is None else np.array(U_init, dtype=np.float64).reshape(T, m).copy()
    params = _f64(params)
    np_ = 0 if params is None else params.size
    X = np.zeros((n_repeat, T + 1, n))
    cost = np.zeros(n_repeat)
    iters = np.zeros(n_repeat, dtype=np.int32)
    rc = lib().ref_ilqr_solve_repeat(model, _p(x0), _p(params), np_, horizon, _p(U), int(n_repeat), int(max_iterations),
                                     ctypes.c_double(tolerance), ctypes.c_double(penalty), int(trig), _p(X), _p(cost), _p(iters, ctypes.c_int))
    if rc:
        raise RuntimeError("ref_ilqr_solve_repeat failed")
    return dict(X=X, U=U, cost=cost, iterations=iters)


def strategy_run_batch(kind, model, x0, params=None, horizon=0, max_outer=10, max_iterations=100, tolerance=1e-5, max_ms=float("inf"),
                       trig=TRIG_GLIBC, threads=0, count_iterations=True):
    """mas::solve(Strategy&, MultiAgentProblem&) per scenario.  x0: [scenarios, agents, n]."""
    x0 = _f64(x0)
    S, A = x0.shape[0], x0.shape[1]
    n, m, T, _ = model_dims(model, horizon)
    params = _f64(params)
    np_ = 0 if params is None else params.shape[-1]
    X = np.zeros((S, A, T + 1, n))
    U = np.zeros((S, A, T, m))
    costs = np.zeros((S, A))
    total = np.zeros(S)
    iters = np.zeros((S, A), dtype=np.int32)
    rc = lib().ref_strategy_run_batch(
        int(kind), model, S, A, _p(x0), _p(params), np_, horizon, int(max_outer), int(max_iterations), ctypes.c_double(tolerance),
        ctypes.c_double(max_ms), int(trig), int(threads), _p(X), _p(U), _p(costs), _p(total),
        _p(iters, ctypes.c_int) if count_iterations else None)
    if rc:
        raise RuntimeError("ref_strategy_run_batch failed")
    return dict(X=X, U=U, costs=costs, total_cost=total, iterations_total=iters)


def _mixed_shapes(models, kind=1):
    dims = [model_dims(m) for m in models]
    if int(kind) == 0:  # centralized: every agent's rows of the stacked solution, horizon of the first block
        dims = [(d[0], d[1], dims[0][2], d[3]) for d in dims]
    return dims, sum(d[0] for d in dims), sum(d[0] * (d[2] + 1) for d in dims), sum(d[1] * d[2] for d in dims)


def strategy_run_mixed(kind, models, x0_list, max_outer=10, max_iterations=100, tolerance=1e-5, trig=TRIG_GLIBC):
    """Strategy (0 centralized, 1 sequential, 2 line search, 3 trust region) over agents of different models.
    x0_list[a]: [scenarios, n_a].  Per-agent lists of X, U back; centralized: iterations_total[:, 0] = stacked iterations."""
    models = [int(m) for m in models]
    A = len(models)
    dims, sx0, sX, sU = _mixed_shapes(models, kind)
    S = np.asarray(x0_list[0]).shape[0]
    x0 = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).reshape(S, -1) for x in x0_list], axis=1))
    X = np.zeros((S, sX))
    U = np.zeros((S, sU))
    costs = np.zeros((S, A))
    total = np.zeros(S)
    iters = np.zeros((S, A), dtype=np.int32)
    marr = np.array(models, dtype=np.int32)
    rc = lib().ref_strategy_run_mixed(int(kind), S, A, _p(marr, ctypes.c_int), _p(x0), int(max_outer), int(max_iterations), ctypes.c_double(tolerance),
                                         int(trig), _p(X), _p(U), _p(costs), _p(total), _p(iters, ctypes.c_int))
    if rc:
        raise RuntimeError("strategy_run_mixed failed")
    Xs, Us, ox, ou = [], [], 0, 0
    for n, m, T, _ in dims:
        Xs.append(X[:, ox:ox + n * (T + 1)].reshape(S, T + 1, n).copy())
        Us.append(U[:, ou:ou + m * T].reshape(S, T, m).copy())
        ox += n * (T + 1)
        ou += m * T
    return dict(X=Xs, U=Us, costs=costs, total_cost=total, iterations_total=iters)


def global_ocp_eval_mixed(models, x0_list, X, U):
    """build_global_ocp of mixed agents (ids = list order, added in reverse), evaluated at (X, U), stage cost at time index 3."""
    models = [int(m) for m in models]
    A = len(models)
    dims, sx0, _, _ = _mixed_shapes(models)
    x0 = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in x0_list]))
    X = _f64(X)
    U = _f64(U)
    dyn = np.zeros(X.size)
    stage, term, dt = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    d4 = np.zeros(4, dtype=np.int32)
    bounds = np.full((2, U.size), np.nan)
    marr = np.array(models, dtype=np.int32)
    rc = lib().ref_global_ocp_eval_mixed(A, _p(marr, ctypes.c_int), _p(x0), _p(X), _p(U), _p(dyn), ctypes.byref(stage), ctypes.byref(term),
                                            _p(d4, ctypes.c_int), ctypes.byref(dt), _p(bounds))
    if rc:
        raise RuntimeError("global_ocp_eval_mixed failed")
    return dict(total_x=int(d4[0]), total_u=int(d4[1]), horizon=int(d4[2]), has_bounds=bool(d4[3]), dt=dt.value, bounds=bounds, dynamics=dyn,
                stage=stage.value, terminal=term.value)
