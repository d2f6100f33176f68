"""GPU: the C++ facade (include/mas_b200/mas_b200.hpp) and the example programs built on it, run as the
reference's binaries would be run, their result line parsed the way the reference's
scripts/compare_solvers.py does (`cost=` / `time_ms=` tokens), and compared with the oracle.
"""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")


@pytest.fixture(scope="module")
def binaries():
    subprocess.check_call(["bash", os.path.join(ROOT, "examples", "build.sh")], stdout=subprocess.DEVNULL)
    return BIN


def run(binaries, name, *args):
    out = subprocess.run([os.path.join(binaries, name), *args], capture_output=True, text=True, timeout=300)
    return out


def parse_line(stdout):
    first = stdout.splitlines()[0]
    fields = dict(tok.split("=", 1) for tok in first.split())
    return fields, first


def parse_block(stdout, label):
    lines = stdout.splitlines()
    i = lines.index(label)
    rows = []
    for ln in lines[i + 2:]:
        if not ln.strip():
            break
        rows.append([float(v) for v in ln.split(",")])
    return lines[i + 1], np.array(rows)


def test_single_track_ocp(binaries, oracle):
    out = run(binaries, "single_track_ocp", "--solver", "ilqr")
    assert out.returncode == 0, out.stderr
    f, line = parse_line(out.stdout)
    assert re.fullmatch(r"solver=ilqr cost=\d+\.\d{6} time_ms=\d+\.\d{6}", line)
    ref = oracle.ilqr_solve_batch(oracle.MODEL_ST_LANE, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iterations=10, tolerance=1e-5, trig=oracle.TRIG_PORTABLE)
    assert abs(float(f["cost"]) - ref["cost"][0]) < 1e-6
    hdr, X = parse_block(out.stdout, "single_track_states")
    assert hdr == "time,x0,x1,x2,x3" and X.shape == (81, 5)
    np.testing.assert_allclose(X[:, 0], np.arange(81) * 0.1, atol=1e-6)
    np.testing.assert_allclose(X[:, 1:], ref["X"][0], atol=1e-5)  # printed with 6 decimals
    hdr, U = parse_block(out.stdout, "single_track_controls")
    assert hdr == "time,u0,u1" and U.shape == (80, 3)
    np.testing.assert_allclose(U[:, 1:], ref["U"][0], atol=1e-5)


def test_rocket_and_pendulum(binaries, oracle):
    out = run(binaries, "rocket_max_altitude", "--solver", "ilqr")
    assert out.returncode == 0, out.stderr
    f, _ = parse_line(out.stdout)
    ref = oracle.ilqr_solve_batch(oracle.MODEL_ROCKET, np.array([[0.0, 0.0, 1.0]]), max_iterations=25, tolerance=1e-6, trig=oracle.TRIG_PORTABLE)
    assert abs(float(f["cost"]) - ref["cost"][0]) < 1e-6 * max(1.0, abs(ref["cost"][0]))
    out = run(binaries, "pendulum_swing_up")
    assert out.returncode == 0, out.stderr
    f, _ = parse_line(out.stdout)
    assert f["solver"] == "ilqr" and np.isfinite(float(f["cost"]))


def test_multi_agent_programs(binaries, oracle):
    out = run(binaries, "multi_agent_lqr", "--agents", "4", "--strategy", "sequential", "--max-outer", "3")
    assert out.returncode == 0, out.stderr
    f, line = parse_line(out.stdout)
    assert line.startswith("solver=ilqr strategy=sequential agents=4 cost=")
    assert abs(float(f["cost"]) - 4 * 20.869847032359) < 1e-5
    assert "agent_3_controls" in out.stdout
    out = run(binaries, "multi_agent_single_track", "3", "--strategy=trust_region_nash", "--max_outer", "10")
    assert out.returncode == 0, out.stderr
    f, _ = parse_line(out.stdout)
    th = 2.0 * np.pi * np.arange(3) / 3
    x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full(3, 4.0)], -1)[None]
    ref = oracle.strategy_run_batch(oracle.STRATEGY_TRUSTREGION, oracle.MODEL_ST_CIRC, x0, max_outer=10, max_iterations=100, tolerance=1e-5,
                                    trig=oracle.TRIG_PORTABLE)
    assert f["strategy"] == "trustregion" and f["agents"] == "3"
    assert abs(float(f["cost"]) - ref["total_cost"][0]) < 1e-6


def test_mixed_agents_through_the_facade(binaries, oracle):
    """Agents of different models in one MultiAgentProblem, C++ facade -> mas_b200_strategy_run_mixed."""
    import math

    th = 0.3 * 2  # libm cos / sin like the C++ builder (the all-FD circular-track agent amplifies a last-bit difference)
    x0 = [np.array([[0.0, 1.0, 0.0, 0.0]]), np.array([[1.0, 0.0, 0.0, 0.0]]),
          np.array([[20.0 * math.cos(th), 20.0 * math.sin(th), 1.57 + th, 4.0]])]
    for name, kind in (("sequential", 1), ("linesearch", 2), ("trustregion", 3)):
        out = run(binaries, "multi_agent_mixed", "--agents", "3", "--strategy", name, "--max-outer", "3")
        assert out.returncode == 0, out.stderr
        f, _ = parse_line(out.stdout)
        ref = oracle.strategy_run_mixed(kind, [0, 2, 1], x0, max_outer=3, max_iterations=8, trig=oracle.TRIG_PORTABLE)
        assert abs(float(f["cost"]) - ref["total_cost"][0]) < 1e-6 * abs(ref["total_cost"][0])
        _, U = parse_block(out.stdout, "agent_1_controls")
        assert U.shape == (10, 5)
    # centralized over the same mix: one stacked solve, horizon of the first agent (80 steps of the lane follower)
    out = run(binaries, "multi_agent_mixed", "--agents", "3", "--strategy", "centralized")
    assert out.returncode == 0, out.stderr
    f, _ = parse_line(out.stdout)
    ref = oracle.strategy_run_mixed(0, [0, 2, 1], x0, max_iterations=8, trig=oracle.TRIG_PORTABLE)
    assert abs(float(f["cost"]) - ref["total_cost"][0]) < 1e-6 * abs(ref["total_cost"][0])
    _, U = parse_block(out.stdout, "agent_1_controls")
    assert U.shape == (80, 5)


def test_error_conventions(binaries):
    out = run(binaries, "single_track_ocp", "--solver", "cgd")
    assert out.returncode == 1 and "Unknown solver 'cgd'" in out.stderr and "Use --help" in out.stderr
    out = run(binaries, "multi_agent_lqr", "--strategy", "bogus")
    assert out.returncode == 1 and "Unknown strategy 'bogus'" in out.stderr
    out = run(binaries, "multi_agent_lqr", "--agents")
    assert out.returncode == 1 and "Missing value" in out.stderr
    out = run(binaries, "multi_agent_lqr", "--help")
    assert out.returncode == 0 and "Available strategies: centralized, sequential, linesearch, trustregion" in out.stdout
