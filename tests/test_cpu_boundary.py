"""CPU-side checks of the drop-in boundary: libmas_b200.so loads, exports every symbol that
include/mas_b200.h declares, host-only entry points work, and device entry points fail loudly (no CPU
fallback) when there is no GPU.  No compute call is made here.
"""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(mas):
    lib = mas.load_library()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mas_b200.h but not exported"


def test_host_only_entry_points(mas):
    for model in range(6):
        info = mas.model_info(model)
        d = mas.example_desc(model)
        assert d.state_dim == info["state_dim"] and d.control_dim == info["control_dim"]
        assert d.deriv_mask == info["example_mask"] and (d.deriv_mask & ~info["available_mask"]) == 0
    d = mas.example_desc(mas.Model.SINGLE_TRACK_LANE)
    assert (d.horizon_steps, d.dt, d.has_input_bounds) == (80, 0.1, 1)
    assert list(d.input_lower)[:2] == [-0.7, -1.0] and list(d.input_upper)[:2] == [0.7, 1.0]
    assert list(d.params)[:5] == [1.0, 10.0, 1.0, 0.1, 0.1]
    with pytest.raises(mas.MasB200Error) as e:
        mas.model_info(99)
    assert e.value.code == 1  # -> std::invalid_argument in the C++ facade
    p = mas.IlqrParams.make(10, 1e-5)
    assert p.penalty == 10.0 and p.penalty_increase == 5.0 and p.constraint_tolerance == 1e-4  # ilqr.hpp:26-37


def test_example_controls_match_oracle(mas, oracle):
    for model in range(6):
        d = mas.example_desc(model)
        assert np.array_equal(mas.example_controls(model, d.horizon_steps), oracle.default_controls(model))


def test_synthetic_inputs_are_deterministic(mas):
    a = mas.synthetic_single_track_x0(1000)
    b = mas.synthetic_single_track_x0(65536)
    assert np.array_equal(a, b[:1000])
    assert np.all(a[:, 0] == 0) and np.all(np.abs(a[:, 1]) <= 2) and np.all(np.abs(a[:, 2]) <= 0.5) and np.all((a[:, 3] >= 0) & (a[:, 3] <= 2))
    # std::mt19937_64(20240607) through std::uniform_real_distribution: first draws are fixed numbers
    np.testing.assert_allclose(a[0], [0.0, 0.32736188, 0.34412183, 0.44811273], atol=1e-8)
    assert not np.array_equal(mas.synthetic_single_track_x0(10, seed=1), a[:10])


def test_no_cpu_fallback(mas):
    """Without a CUDA device every context-touching call fails with MAS_B200_ERR_CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    with pytest.raises(mas.MasB200Error) as e:
        mas.Context(0)
    assert e.value.code == 3


def test_oracle_is_not_reachable_from_the_product():
    """The product never imports, links or names the oracle."""
    pkg = os.path.join(ROOT, "multi_agent_solver_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.lower(), (base, f)
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(base, f), errors="ignore").read().lower()
            assert "oracle/" not in text and "import oracle" not in text
    lib = os.path.join(pkg, "libmas_b200.so")
    needed = os.popen(f"objdump -p {lib} | grep NEEDED").read()
    assert "oracle" not in needed


def test_mixed_agent_structure_without_a_device(mas):
    """mas_b200_global_ocp_eval_mixed, structure only (X = U = NULL needs no device): the reference's
    MultiAgentProblemTest.BuildGlobalProblemMergesAgents checks (tests/ocp_tests.cpp:76-126) on registered models of
    different shapes -- id-sorted blocks, offsets, summed dims, horizon / dt of the first block, bounds only if all have them."""
    pend, rocket, lqr = mas.example_desc(3), mas.example_desc(4), mas.example_desc(2)
    got = mas.global_ocp_eval_mixed(None, [rocket, pend], agent_ids=[2, 1])
    assert list(got["block_agent"]) == [1, 0]  # the agent with id 1 (pendulum, 2 x 1) comes first
    assert (got["total_x"], got["total_u"]) == (2 + 3, 1 + 1)
    assert list(got["state_offsets"]) == [0, 2] and list(got["control_offsets"]) == [0, 1]
    assert got["horizon"] == pend.horizon_steps and got["dt"] == pend.dt
    assert got["has_bounds"] and list(got["bounds"][0]) == [-5.0, 0.0] and list(got["bounds"][1]) == [5.0, 20.0]
    got = mas.global_ocp_eval_mixed(None, [rocket, pend, lqr], agent_ids=[2, 1, 0])
    assert not got["has_bounds"] and got["total_x"] == 9 and got["horizon"] == lqr.horizon_steps  # LQR has no bounds and is block 0


def test_cpp_facade_host_behaviour(mas):
    """include/mas_b200/mas_b200.hpp without a device (tests/csrc/facade_host_test.cpp): set_params' std::out_of_range on a
    missing key (ilqr.hpp:42-44), the name registries' std::invalid_argument (example_utils.hpp:32-110), column-major
    Matrix, verify_problem, the OCP description handed to the C ABI, compute_offsets on out-of-order ids the way the
    reference's tests/ocp_tests.cpp:76-154 checks it, and -- here, where there is no GPU -- std::runtime_error from the
    first call that needs the device (no host implementation behind the facade)."""
    import subprocess

    src = os.path.join(ROOT, "tests", "csrc", "facade_host_test.cpp")
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "facade_host_test")
    lib_dir = os.path.join(ROOT, "multi_agent_solver_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", exe, "-L" + lib_dir,
                           "-lmas_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ALL OK" in out.stdout, out.stdout + out.stderr
