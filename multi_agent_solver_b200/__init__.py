"""multi_agent_solver_b200 -- B200-native batched iLQR engine (drop-in for the iLQR path of
markomiz/multi_agent_solver).

The product is `libmas_b200.so` (hand-written sm_100a CUDA behind the C ABI of `include/mas_b200.h`)
plus the C++ facade in `include/mas_b200/*.hpp`.  This Python package is only the ctypes binding the
tests and `bench.py` use to call that C ABI; it holds no algorithm and has no CPU fallback: importing
`capi` without the built library raises, and every call needs a CUDA device.
"""
from .capi import (  # noqa: F401
    Batch,
    Context,
    DerivBits,
    IlqrParams,
    MasB200Error,
    Model,
    OcpDesc,
    Status,
    Strategy,
    example_controls,
    example_desc,
    ilqr_solve_batch,
    library_path,
    load_library,
    model_info,
    global_ocp_eval_mixed,
    strategy_run,
    strategy_run_mixed,
    synthetic_single_track_x0,
)
