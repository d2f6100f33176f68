"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes front-end to oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Parity pin: oracle == oracle/_ref/libref.so (the reference's own sources) bit for bit, tests/test_ref_pin.py;
see oracle/dense.hpp.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

MODEL_ST_LANE, MODEL_ST_CIRC, MODEL_LQR, MODEL_PENDULUM, MODEL_ROCKET, MODEL_ST_LANE_CON = range(6)
STRATEGY_CENTRALIZED, STRATEGY_SEQUENTIAL, STRATEGY_LINESEARCH, STRATEGY_TRUSTREGION = range(4)
TRIG_GLIBC, TRIG_PORTABLE = 0, 1
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_TIME_LIMIT = 0, 1, 2

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (g++ only, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "dense.hpp", "ref_core.hpp", "ref_ilqr.hpp", "ref_models.hpp", "ref_multi_agent.hpp")]
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "all"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a, ct=ctypes.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def model_dims(model: int, horizon: int = 0):
    n, m, T, dt = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    rc = lib().oracle_model_dims(model, horizon, ctypes.byref(n), ctypes.byref(m), ctypes.byref(T), ctypes.byref(dt))
    if rc:
        raise ValueError(f"oracle: unknown model {model}")
    return n.value, m.value, T.value, dt.value


def max_threads() -> int:
    return lib().oracle_max_threads()


def synthetic_single_track_x0(batch: int, seed: int = 20240607) -> np.ndarray:
    """Config-3 initial states: std::mt19937_64(seed); Y, psi, v per problem."""
    x0 = np.empty((batch, 4))
    lib().oracle_synthetic_single_track_x0(ctypes.c_ulonglong(seed), int(batch), _p(x0))
    return x0


def default_controls(model: int, horizon: int = 0) -> np.ndarray:
    n, m, T, _ = model_dims(model, horizon)
    U = np.zeros((T, m))
    if lib().oracle_default_controls(model, horizon, _p(U)):
        raise RuntimeError("oracle_default_controls failed")
    return U


def rollout_cost(model, x0, U, params=None, horizon=0, trig=TRIG_GLIBC):
    x0 = _f64(x0)
    U = _f64(U)
    batch = x0.shape[0]
    n, m, T, _ = model_dims(model, horizon)
    params = _f64(params)
    np_ = 0 if params is None else params.shape[1]
    X = np.zeros((batch, T + 1, n))
    cost = np.zeros(batch)
    rc = lib().oracle_rollout_cost(model, batch, _p(x0), _p(params), np_, horizon, _p(U), trig, _p(X), _p(cost))
    if rc:
        raise RuntimeError("oracle_rollout_cost failed")
    return X, cost


def ilqr_solve_batch(model, x0, U_init=None, params=None, horizon=0, max_iterations=10, tolerance=1e-5, max_ms=float("inf"),
                     trig=TRIG_GLIBC, aliased_sym=True, threads=0):
    """Returns dict(X[batch,T+1,n], U[batch,T,m], cost, iterations, status, rollouts, alpha_trials, reg_retries)."""
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
    iters = np.zeros(batch, dtype=np.int32)
    status = np.zeros(batch, dtype=np.int32)
    stats = np.zeros((batch, 3), dtype=np.int32)
    rc = lib().oracle_ilqr_solve_batch(
        model, batch, _p(x0), _p(params), np_, horizon, _p(U), int(max_iterations), ctypes.c_double(tolerance), ctypes.c_double(max_ms),
        int(trig), int(bool(aliased_sym)), int(threads), _p(X), _p(cost), _p(iters, ctypes.c_int), _p(status, ctypes.c_int),
        _p(stats, ctypes.c_int))
    if rc:
        raise RuntimeError("oracle_ilqr_solve_batch failed")
    return dict(X=X, U=U, cost=cost, iterations=iters, status=status, rollouts=stats[:, 0], alpha_trials=stats[:, 1], reg_retries=stats[:, 2])


def ilqr_solve_repeat(model, x0, n_repeat, U_init=None, params=None, horizon=0, max_iterations=10, tolerance=1e-5, penalty=10.0,
                      trig=TRIG_GLIBC):
    """One solver object, n_repeat solve() calls on the same OCP (multipliers / penalty persist).  Per-solve outputs."""
    x0 = _f64(x0).reshape(-1)
    n, m, T, _ = model_dims(model, horizon)
    U = default_controls(model, horizon) if U_init is None else np.array(U_init, dtype=np.float64).reshape(T, m).copy()
    params = _f64(params)
    np_ = 0 if params is None else params.size
    X = np.zeros((n_repeat, T + 1, n))
    cost = np.zeros(n_repeat)
    iters = np.zeros(n_repeat, dtype=np.int32)
    status = np.zeros(n_repeat, dtype=np.int32)
    rc = lib().oracle_ilqr_solve_repeat(model, _p(x0), _p(params), np_, horizon, _p(U), int(n_repeat), int(max_iterations),
                                        ctypes.c_double(tolerance), ctypes.c_double(penalty), int(trig), _p(X), _p(cost), _p(iters, ctypes.c_int),
                                        _p(status, ctypes.c_int))
    if rc:
        raise RuntimeError("oracle_ilqr_solve_repeat failed")
    return dict(X=X, U=U, cost=cost, iterations=iters, status=status)


def ilqr_solve_trace(model, x0, U_init=None, params=None, horizon=0, max_iterations=10, tolerance=1e-5, trig=TRIG_GLIBC, aliased_sym=True):
    x0 = _f64(x0).reshape(-1)
    n, m, T, _ = model_dims(model, horizon)
    U = default_controls(model, horizon) if U_init is None else np.array(U_init, dtype=np.float64).reshape(T, m).copy()
    params = _f64(params)
    np_ = 0 if params is None else params.size
    cost_trace = np.zeros(max_iterations)
    alpha_trace = np.zeros(max_iterations, dtype=np.int32)
    n_trace = ctypes.c_int()
    stats = np.zeros(3, dtype=np.int32)
    rc = lib().oracle_ilqr_solve_trace(model, _p(x0), _p(params), np_, horizon, _p(U), int(max_iterations), ctypes.c_double(tolerance), int(trig),
                                       int(bool(aliased_sym)), _p(cost_trace), _p(alpha_trace, ctypes.c_int), ctypes.byref(n_trace),
                                       _p(stats, ctypes.c_int))
    if rc:
        raise RuntimeError("oracle_ilqr_solve_trace failed")
    k = n_trace.value
    return dict(U=U, cost_trace=cost_trace[:k], alpha_index=alpha_trace[:k], rollouts=int(stats[0]), alpha_trials=int(stats[1]),
                reg_retries=int(stats[2]))


def strategy_run_batch(kind, model, x0, params=None, horizon=0, max_outer=10, max_iterations=100, tolerance=1e-5, max_ms=float("inf"),
                       trig=TRIG_GLIBC, aliased_sym=True, threads=0):
    """x0: [scenarios, agents, n].  Returns dict(X, U, costs, total_cost, trace_iters, trace_accept, trace_cost)."""
    x0 = _f64(x0)
    S, A = x0.shape[0], x0.shape[1]
    n, m, T, _ = model_dims(model, horizon)
    params = _f64(params)
    np_ = 0 if params is None else params.shape[-1]
    X = np.zeros((S, A, T + 1, n))
    U = np.zeros((S, A, T, m))
    costs = np.zeros((S, A))
    total = np.zeros(S)
    t_it = np.zeros((S, max_outer, A), dtype=np.int32)
    t_acc = np.zeros((S, max_outer, A), dtype=np.int32)
    t_cost = np.zeros((S, max_outer, A))
    rc = lib().oracle_strategy_run_batch(
        int(kind), model, S, A, _p(x0), _p(params), np_, horizon, int(max_outer), int(max_iterations), ctypes.c_double(tolerance),
        ctypes.c_double(max_ms), int(trig), int(bool(aliased_sym)), int(threads), _p(X), _p(U), _p(costs), _p(total), _p(t_it, ctypes.c_int),
        _p(t_acc, ctypes.c_int), _p(t_cost))
    if rc:
        raise RuntimeError("oracle_strategy_run_batch failed")
    return dict(X=X, U=U, costs=costs, total_cost=total, trace_iters=t_it, trace_accept=t_acc, trace_cost=t_cost)


def global_ocp_eval(model, x0, X, U, params=None, horizon=0):
    x0 = _f64(x0)
    A = x0.shape[0]
    params = _f64(params)
    np_ = 0 if params is None else params.shape[-1]
    X = _f64(X)
    U = _f64(U)
    dyn = np.zeros(X.size)
    stage = ctypes.c_double()
    term = ctypes.c_double()
    dims = np.zeros(3, dtype=np.int32)
    rc = lib().oracle_global_ocp_eval(model, A, _p(x0), _p(params), np_, horizon, _p(X), _p(U), _p(dyn), ctypes.byref(stage), ctypes.byref(term),
                                      _p(dims, ctypes.c_int))
    if rc:
        raise RuntimeError("oracle_global_ocp_eval failed")
    return dyn, stage.value, term.value, dims


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
    rc = lib().oracle_strategy_run_mixed(int(kind), S, A, _p(marr, ctypes.c_int), _p(x0), int(max_outer), int(max_iterations), ctypes.c_double(tolerance),
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
    rc = lib().oracle_global_ocp_eval_mixed(A, _p(marr, ctypes.c_int), _p(x0), _p(X), _p(U), _p(dyn), ctypes.byref(stage), ctypes.byref(term),
                                            _p(d4, ctypes.c_int), ctypes.byref(dt), _p(bounds))
    if rc:
        raise RuntimeError("global_ocp_eval_mixed failed")
    return dict(total_x=int(d4[0]), total_u=int(d4[1]), horizon=int(d4[2]), has_bounds=bool(d4[3]), dt=dt.value, bounds=bounds, dynamics=dyn,
                stage=stage.value, terminal=term.value)
