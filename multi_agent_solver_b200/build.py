"""Builds multi_agent_solver_b200/libmas_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree on purpose: the .so travels with the repository snapshot to the GPU box.  Each model is its
own translation unit so the files compile in parallel.  `-fmad=false`: the kernels restate the
reference's arithmetic without fused multiply-add (the reference Release build targets baseline
x86-64); the only fma instructions come from the explicit fma() calls of portable_math.h.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
# MAS_B200_VARIANT=<name>: a tuning build next to the product library (tools/_variants/libmas_b200_<name>.so, own object
# directory), selected at run time with MAS_B200_LIB=<path>; used for A/B measurements only.
VARIANT = os.environ.get("MAS_B200_VARIANT", "")
OBJ_DIR = os.path.join(PKG_DIR, "_obj" + ("_" + VARIANT if VARIANT else ""))
LIB_PATH = os.path.join(PKG_DIR, "libmas_b200.so") if not VARIANT else os.path.join(ROOT, "tools", "_variants", f"libmas_b200_{VARIANT}.so")

SOURCES = ["engine.cu", "capi.cu", "centralized.cu", "model_st_lane.cu", "model_st_lane_con.cu", "model_st_circ.cu", "model_lqr4.cu", "model_pendulum.cu", "model_rocket.cu"]
HEADERS = ["engine.cuh", "ilqr_core.cuh", "models.cuh", "centralized.cuh", "centralized_host.cuh", "stacked_mixed.cuh", "stacked_mixed_host.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
] + os.environ.get("MAS_B200_EXTRA_NVCC_FLAGS", "").split()  # tuning experiments, e.g. -DMAS_MIN_CTAS=9


def _newest_input() -> float:
    paths = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    paths += [os.path.join(ROOT, "include", "mas_b200.h"), os.path.join(ROOT, "include", "mas_b200", "portable_math.h"), __file__]
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_input():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        objs = list(pool.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-ccbin", "/usr/bin/g++", "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(f"built {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
