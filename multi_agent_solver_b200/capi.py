"""ctypes binding of include/mas_b200.h (libmas_b200.so).  No algorithm lives here."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("MAS_B200_LIB") or os.path.join(_PKG, "libmas_b200.so")  # MAS_B200_LIB: a tuning build (tools/_variants)
_lib: Optional[ctypes.CDLL] = None

MAX_CONTROL_DIM = 8
MAX_PARAMS = 8


class Model:
    SINGLE_TRACK_LANE, SINGLE_TRACK_CIRC, LQR4, PENDULUM, ROCKET, SINGLE_TRACK_LANE_CONSTRAINED = range(6)


class Status:
    CONVERGED, MAX_ITER, TIME_LIMIT = range(3)


class Strategy:
    CENTRALIZED, SEQUENTIAL, LINESEARCH, TRUSTREGION = range(4)


class DerivBits:
    A, B, LX, LU, LXX, LUU, LUX, VX, VXX, EQ_JX, EQ_JU, INEQ_JX, INEQ_JU = (1 << i for i in range(13))


class MasB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mas_b200 error {code}: {msg}")
        self.code = code


class OcpDesc(ctypes.Structure):
    _fields_ = [
        ("model_id", ctypes.c_int),
        ("state_dim", ctypes.c_int),
        ("control_dim", ctypes.c_int),
        ("horizon_steps", ctypes.c_int),
        ("dt", ctypes.c_double),
        ("deriv_mask", ctypes.c_uint),
        ("has_input_bounds", ctypes.c_int),
        ("input_lower", ctypes.c_double * MAX_CONTROL_DIM),
        ("input_upper", ctypes.c_double * MAX_CONTROL_DIM),
        ("num_params", ctypes.c_int),
        ("params", ctypes.c_double * MAX_PARAMS),
    ]


class IlqrParams(ctypes.Structure):
    _fields_ = [
        ("max_iterations", ctypes.c_int),
        ("tolerance", ctypes.c_double),
        ("max_ms", ctypes.c_double),
        ("debug", ctypes.c_int),
        ("penalty", ctypes.c_double),
        ("penalty_increase", ctypes.c_double),
        ("constraint_tolerance", ctypes.c_double),
        ("inequality_activation_tolerance", ctypes.c_double),
    ]

    @staticmethod
    def make(max_iterations: int, tolerance: float, max_ms: float = float("inf")) -> "IlqrParams":
        p = IlqrParams()
        load_library().mas_b200_ilqr_default_params(ctypes.byref(p))
        p.max_iterations = int(max_iterations)
        p.tolerance = float(tolerance)
        p.max_ms = float(max_ms)
        return p


class DeviceView(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int), ("ld", ctypes.c_int), ("state_dim", ctypes.c_int), ("control_dim", ctypes.c_int), ("horizon_steps", ctypes.c_int),
        ("x0", ctypes.c_void_p), ("X", ctypes.c_void_p), ("U", ctypes.c_void_p), ("cost", ctypes.c_void_p),
        ("iterations", ctypes.c_void_p), ("status", ctypes.c_void_p), ("params", ctypes.c_void_p),
    ]


class BatchStats(ctypes.Structure):
    _fields_ = [
        ("iterations", ctypes.c_longlong), ("alpha_trials", ctypes.c_longlong), ("reg_retries", ctypes.c_longlong),
        ("kernel_launches", ctypes.c_longlong), ("outer_iterations_run", ctypes.c_int), ("forward_lanes", ctypes.c_int),
        ("forward_chains", ctypes.c_int),
    ]


class Profile(ctypes.Structure):
    _fields_ = [
        ("prologue_ms", ctypes.c_double), ("backward_ms", ctypes.c_double), ("forward_ms", ctypes.c_double),
        ("prologue_launches", ctypes.c_longlong), ("backward_launches", ctypes.c_longlong), ("forward_launches", ctypes.c_longlong),
        ("problem_iterations", ctypes.c_longlong), ("solves", ctypes.c_longlong),
    ]


def library_path() -> str:
    return _LIB_PATH


def load_library() -> ctypes.CDLL:
    """Loads libmas_b200.so.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `python -m multi_agent_solver_b200.build` (needs nvcc). "
                              "multi_agent_solver_b200 has no CPU fallback.")
        lib = ctypes.CDLL(_LIB_PATH)
        lib.mas_b200_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise MasB200Error(rc, load_library().mas_b200_last_error().decode())


def _dptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _iptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def model_info(model_id: int):
    nx, nu, npar = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    avail, ex = ctypes.c_uint(), ctypes.c_uint()
    dp = (ctypes.c_double * MAX_PARAMS)()
    _check(load_library().mas_b200_model_info(model_id, ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(npar), ctypes.byref(avail),
                                               ctypes.byref(ex), dp))
    return dict(state_dim=nx.value, control_dim=nu.value, num_params=npar.value, available_mask=avail.value, example_mask=ex.value,
                default_params=list(dp)[: npar.value])


def example_desc(model_id: int, horizon_steps: int = 0) -> OcpDesc:
    d = OcpDesc()
    _check(load_library().mas_b200_example_desc(model_id, ctypes.byref(d)))
    if horizon_steps > 0:
        d.horizon_steps = horizon_steps
        if model_id == Model.PENDULUM:
            d.params[0] = float(horizon_steps)
    return d


def synthetic_single_track_x0(batch: int, seed: int = 20240607) -> np.ndarray:
    """Config-3 initial states: std::mt19937_64(seed), Y, psi, v per problem."""
    x0 = np.empty((batch, 4))
    _check(load_library().mas_b200_synthetic_single_track_x0(ctypes.c_ulonglong(seed), int(batch), _dptr(x0)))
    return x0


def example_controls(model_id: int, horizon_steps: int) -> np.ndarray:
    nu = model_info(model_id)["control_dim"]
    U = np.zeros((horizon_steps, nu))
    _check(load_library().mas_b200_example_controls(model_id, horizon_steps, _dptr(U)))
    return U


class Context:
    """mas_b200_context_t: one device, one stream."""

    def __init__(self, device_id: int = -1, stream: int = 0):
        self._h = ctypes.c_void_p()
        _check(load_library().mas_b200_context_create(int(device_id), ctypes.c_void_p(stream or None), ctypes.byref(self._h)))

    def synchronize(self) -> None:
        _check(load_library().mas_b200_context_synchronize(self._h))

    def init_nccl(self, unique_id: bytes, rank: int, world_size: int) -> None:
        buf = ctypes.create_string_buffer(unique_id, 128)
        _check(load_library().mas_b200_context_init_nccl(self._h, buf, int(rank), int(world_size)))

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(load_library().mas_b200_nccl_unique_id(buf))
        return buf.raw

    def set_blocking_sync(self, enable: bool) -> None:
        _check(load_library().mas_b200_context_set_blocking_sync(self._h, int(bool(enable))))

    def set_agent_sharding(self, agents_sharded: bool) -> None:
        _check(load_library().mas_b200_context_set_agent_sharding(self._h, int(bool(agents_sharded))))

    def strategy_joint(self, world: int, n_scenarios: int, n_agents: int, desc: "OcpDesc") -> dict:
        """mas_b200_strategy_get_joint: every rank's agents after the last round, [world, scenarios, agents, ...]."""
        T, nx, nu = desc.horizon_steps, desc.state_dim, desc.control_dim
        X = np.empty((world, n_scenarios, n_agents, T + 1, nx))
        U = np.empty((world, n_scenarios, n_agents, T, nu))
        costs = np.empty((world, n_scenarios, n_agents))
        _check(load_library().mas_b200_strategy_get_joint(self._h, _dptr(X), _dptr(U), _dptr(costs)))
        return dict(X=X, U=U, costs=costs)

    def exchange_stats(self) -> dict:
        ms, rounds, nbytes = ctypes.c_double(), ctypes.c_int(), ctypes.c_longlong()
        _check(load_library().mas_b200_strategy_get_exchange_stats(self._h, ctypes.byref(ms), ctypes.byref(rounds), ctypes.byref(nbytes)))
        return dict(collective_ms=ms.value, rounds=rounds.value, bytes_per_round=nbytes.value)

    def probe_fp64_peak(self) -> float:
        tf = ctypes.c_double()
        _check(load_library().mas_b200_probe_fp64_peak(self._h, ctypes.byref(tf)))
        return tf.value

    def selftest_division(self, pairs: int, seed: int = 1) -> dict:
        counts = (ctypes.c_longlong * 5)()
        _check(load_library().mas_b200_selftest_division(self._h, ctypes.c_ulonglong(seed), ctypes.c_longlong(pairs), counts))
        return dict(checked=counts[0], div_exact=counts[1], div_mismatch=counts[2], div_const_exact=counts[3], div_const_mismatch=counts[4])

    def close(self) -> None:
        if self._h:
            load_library().mas_b200_context_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """mas_b200_batch_t: `batch` same-shaped OCPs resident in HBM."""

    def __init__(self, ctx: Context, desc: OcpDesc, batch: int):
        self.ctx = ctx
        self.desc = desc
        self.batch = int(batch)
        self.nx, self.nu, self.T = desc.state_dim, desc.control_dim, desc.horizon_steps
        self._h = ctypes.c_void_p()
        _check(load_library().mas_b200_batch_create(ctx._h, ctypes.byref(desc), self.batch, ctypes.byref(self._h)))

    def set_initial_states(self, x0) -> None:
        x0 = _f64(x0)
        assert x0.shape == (self.batch, self.nx)
        self._x0_keep = x0
        _check(load_library().mas_b200_batch_set_initial_states(self._h, _dptr(x0)))

    def set_params(self, params) -> None:
        params = _f64(params)
        self._p_keep = params
        _check(load_library().mas_b200_batch_set_params(self._h, _dptr(params)))

    def set_controls(self, U=None) -> None:
        U = _f64(U)
        if U is not None:
            assert U.shape == (self.batch, self.T, self.nu)
        self._u_keep = U
        _check(load_library().mas_b200_batch_set_controls(self._h, _dptr(U)))

    def initialize(self) -> None:
        _check(load_library().mas_b200_batch_initialize(self._h))

    def solve(self, params: IlqrParams) -> None:
        _check(load_library().mas_b200_batch_solve(self._h, ctypes.byref(params)))

    def set_trial_store(self, enable: bool) -> None:
        _check(load_library().mas_b200_batch_set_trial_store(self._h, int(enable)))

    def reset_solver_state(self) -> None:
        _check(load_library().mas_b200_batch_reset_solver_state(self._h))

    def set_profiling(self, enable: bool) -> None:
        _check(load_library().mas_b200_batch_set_profiling(self._h, int(bool(enable))))

    def profile(self) -> dict:
        p = Profile()
        _check(load_library().mas_b200_batch_get_profile(self._h, ctypes.byref(p)))
        return {k: getattr(p, k) for k, _ in Profile._fields_}

    def set_tuning(self, forward_lanes: int = 0, forward_chains: int = 0) -> None:
        _check(load_library().mas_b200_batch_set_tuning(self._h, int(forward_lanes), int(forward_chains)))

    def set_backward_mode(self, mode: int, max_problems: int = 0) -> None:
        """0 auto, 1 one thread per problem, 2 FD tasks over eight lanes, 3 time-parallel linearisation + Riccati sweep."""
        _check(load_library().mas_b200_batch_set_backward_mode(self._h, int(mode), int(max_problems)))

    def debug_trace(self, problem: int, max_records: int = 4096) -> np.ndarray:
        """mas_b200_batch_get_debug_trace: [records, 6] = cost, merit, d_merit, eq_violation, ineq_violation, accepted index."""
        rec = np.empty((max_records, 6))
        n = ctypes.c_int()
        _check(load_library().mas_b200_batch_get_debug_trace(self._h, int(problem), int(max_records), _dptr(rec), ctypes.byref(n)))
        return rec[: n.value].copy()

    def set_concurrency_hint(self, solves_in_flight: int) -> None:
        _check(load_library().mas_b200_batch_set_concurrency_hint(self._h, int(solves_in_flight)))

    def set_line_search_mode(self, mode: int) -> None:
        """0 auto, 1 concurrent lanes, 2 compacted rounds."""
        _check(load_library().mas_b200_batch_set_line_search_mode(self._h, int(mode)))

    def get_solution(self, out=None):
        if out is None:
            out = dict(X=np.empty((self.batch, self.T + 1, self.nx)), U=np.empty((self.batch, self.T, self.nu)), cost=np.empty(self.batch),
                       iterations=np.empty(self.batch, dtype=np.int32), status=np.empty(self.batch, dtype=np.int32))
        _check(load_library().mas_b200_batch_get_solution(self._h, _dptr(out.get("X")), _dptr(out.get("U")), _dptr(out.get("cost")),
                                                           _iptr(out.get("iterations")), _iptr(out.get("status"))))
        return out

    def begin_get_solution(self, out) -> None:
        """mas_b200_batch_begin_get_solution: asynchronous download into `out` (pinned arrays); pair with wait_solution()."""
        self._out_keep = out
        _check(load_library().mas_b200_batch_begin_get_solution(self._h, _dptr(out.get("X")), _dptr(out.get("U")), _dptr(out.get("cost")),
                                                                 _iptr(out.get("iterations")), _iptr(out.get("status"))))

    def wait_solution(self) -> None:
        _check(load_library().mas_b200_batch_wait_solution(self._h))

    def set_result_sink(self, out) -> None:
        """mas_b200_batch_set_result_sink: every following solve streams its results into `out` (page-locked arrays, e.g.
        torch pin_memory) while it runs; wait_solution() fences them.  None unregisters."""
        out = out or {}
        self._sink_keep = out
        _check(load_library().mas_b200_batch_set_result_sink(self._h, _dptr(out.get("X")), _dptr(out.get("U")), _dptr(out.get("cost")),
                                                              _iptr(out.get("iterations")), _iptr(out.get("status"))))

    def device_view(self) -> DeviceView:
        v = DeviceView()
        _check(load_library().mas_b200_batch_get_device_view(self._h, ctypes.byref(v)))
        return v

    def stats(self) -> dict:
        s = BatchStats()
        _check(load_library().mas_b200_batch_get_stats(self._h, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in BatchStats._fields_}

    def close(self) -> None:
        if self._h:
            load_library().mas_b200_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ilqr_solve_batch(ctx: Context, desc: OcpDesc, params: IlqrParams, x0, U=None, model_params=None):
    """mas_b200_ilqr_solve_batch: one-shot solve on host arrays."""
    x0 = _f64(x0)
    batch = x0.shape[0]
    T, nx, nu = desc.horizon_steps, desc.state_dim, desc.control_dim
    Uio = None if U is None else np.array(U, dtype=np.float64, order="C").reshape(batch, T, nu).copy()
    model_params = _f64(model_params)
    X = np.empty((batch, T + 1, nx))
    cost = np.empty(batch)
    it = np.empty(batch, dtype=np.int32)
    st = np.empty(batch, dtype=np.int32)
    if Uio is None:
        Uio = np.zeros((batch, T, nu))
    _check(load_library().mas_b200_ilqr_solve_batch(ctx._h, ctypes.byref(desc), ctypes.byref(params), batch, _dptr(x0), _dptr(model_params),
                                                     _dptr(Uio), _dptr(X), _dptr(cost), _iptr(it), _iptr(st)))
    return dict(X=X, U=Uio, cost=cost, iterations=it, status=st)


def strategy_run(ctx: Context, strategy: int, desc: OcpDesc, params: IlqrParams, max_outer: int, x0, model_params=None, trace: bool = True,
                 U_init=None):
    """mas_b200_strategy_run.  x0: [scenarios, agents, n]."""
    x0 = _f64(x0)
    S, A = x0.shape[0], x0.shape[1]
    T, nx, nu = desc.horizon_steps, desc.state_dim, desc.control_dim
    model_params = _f64(model_params)
    U_init = _f64(U_init)
    X = np.empty((S, A, T + 1, nx))
    U = np.empty((S, A, T, nu))
    costs = np.empty((S, A))
    total = np.empty(S)
    t_it = np.zeros((S, max_outer, A), dtype=np.int32) if trace else None
    t_acc = np.zeros((S, max_outer, A), dtype=np.int32) if trace else None
    t_cost = np.zeros((S, max_outer, A)) if trace else None
    _check(load_library().mas_b200_strategy_run(ctx._h, int(strategy), ctypes.byref(desc), ctypes.byref(params), int(max_outer), S, A, _dptr(x0),
                                                 _dptr(model_params), _dptr(U_init), _dptr(X), _dptr(U), _dptr(costs), _dptr(total), _iptr(t_it), _iptr(t_acc),
                                                 _dptr(t_cost)))
    return dict(X=X, U=U, costs=costs, total_cost=total, trace_iters=t_it, trace_accept=t_acc, trace_cost=t_cost)


def strategy_run_mixed(ctx: Context, strategy: int, descs, params: IlqrParams, max_outer: int, x0_list, model_params=None, U_init=None):
    """mas_b200_strategy_run_mixed.  descs: one OcpDesc per agent; x0_list[a]: [scenarios, n_a]."""
    A = len(descs)
    x0_list = [_f64(x) for x in x0_list]
    S = x0_list[0].shape[0]
    darr = (OcpDesc * A)(*descs)
    PD = ctypes.POINTER(ctypes.c_double)

    def ptrs(arrs):
        return (PD * A)(*[(_dptr(a) if a is not None else PD()) for a in arrs])

    # the centralized strategy returns every agent's rows of the stacked solution: horizon of the first agent
    T_of = (lambda d: descs[0].horizon_steps) if int(strategy) == int(Strategy.CENTRALIZED) else (lambda d: d.horizon_steps)
    X = [np.empty((S, T_of(d) + 1, d.state_dim)) for d in descs]
    U = [np.empty((S, T_of(d), d.control_dim)) for d in descs]
    costs = [np.empty(S) for _ in descs]
    total = np.empty(S)
    t_it = np.zeros((S, max_outer, A), dtype=np.int32)
    mp = [None] * A if model_params is None else [_f64(m) for m in model_params]
    u0 = [None] * A if U_init is None else [_f64(u) for u in U_init]
    _check(load_library().mas_b200_strategy_run_mixed(ctx._h, int(strategy), darr, ctypes.byref(params), int(max_outer), S, A, ptrs(x0_list), ptrs(mp),
                                                       ptrs(u0), ptrs(X), ptrs(U), ptrs(costs), _dptr(total), _iptr(t_it)))
    return dict(X=X, U=U, costs=np.stack(costs, -1), total_cost=total, trace_iters=t_it)


def global_ocp_eval_mixed(ctx: Context, descs, agent_ids=None, X=None, U=None, time_index: int = 0):
    """mas_b200_global_ocp_eval_mixed: structure (and, with X / U, values) of the stacked problem of mixed agents."""
    A = len(descs)
    darr = (OcpDesc * A)(*descs)
    ids = None if agent_ids is None else (ctypes.c_ulonglong * A)(*[int(i) for i in agent_ids])
    dims = np.zeros(4, dtype=np.int32)
    dt = ctypes.c_double()
    total_u = sum(d.control_dim for d in descs)
    total_x = sum(d.state_dim for d in descs)
    bounds = np.full((2, total_u), np.nan)
    block_agent = np.zeros(A, dtype=np.int32)
    soff = np.zeros(A, dtype=np.int32)
    uoff = np.zeros(A, dtype=np.int32)
    X = _f64(X)
    U = _f64(U)
    dyn = np.zeros(total_x)
    stage, term = ctypes.c_double(), ctypes.c_double()
    _check(load_library().mas_b200_global_ocp_eval_mixed(ctx._h if ctx is not None else None, darr, ids, A, _dptr(X), _dptr(U), int(time_index), _dptr(dyn), ctypes.byref(stage),
                                                          ctypes.byref(term), _iptr(dims), ctypes.byref(dt), _dptr(bounds), _iptr(block_agent),
                                                          _iptr(soff), _iptr(uoff)))
    return dict(total_x=int(dims[0]), total_u=int(dims[1]), horizon=int(dims[2]), has_bounds=bool(dims[3]), dt=dt.value, bounds=bounds,
                block_agent=block_agent, state_offsets=soff, control_offsets=uoff, dynamics=dyn, stage=stage.value, terminal=term.value)
