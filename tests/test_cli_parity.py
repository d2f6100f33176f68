"""The example programs' command line against the reference's OWN main()s, on command lines that need no device.

oracle/_ref/libref.so holds the reference's five example sources unmodified, their `main` renamed by the preprocessor
(oracle/ref/ex_*.cpp), so the reference's argument parser (examples/cli.hpp), usage text and error messages can be run
here and compared byte for byte with examples/bin/* (CLI + result-line format).  The one intended
difference is the solver list: `cgd` is not on the device path.  Second half: whole runs of the reference's mains
(`--solver ilqr`, every strategy) -- what they print (result line, CSV blocks) is what the oracle computes with the
parameters of SURVEY 8a's config table, to the printed precision; tests/test_gpu_facade.py compares examples/bin/* with
the same oracle on the GPU.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "bin")

# `int main(int, char**)` of each example, renamed to ref_example_main_<name> (Itanium mangling of (int, char**))
REF_MAIN = {name: f"_Z{len('ref_example_main_' + name)}ref_example_main_{name}iPPc"
            for name in ("single_track_ocp", "multi_agent_single_track", "multi_agent_lqr", "pendulum_swing_up", "rocket_max_altitude")}


@pytest.fixture(scope="module")
def ref_lib():
    from oracle import ref_py

    if not ref_py.available():
        pytest.skip("oracle/_ref/libref.so absent and no reference sources to build it from")
    return ref_py.build()


@pytest.fixture(scope="module")
def binaries(mas):
    subprocess.check_call(["bash", os.path.join(ROOT, "examples", "build.sh")], stdout=subprocess.DEVNULL)
    return BIN


def run_reference_main(lib_path, name, args):
    """The reference's main() in a child process (it writes to the process's stdout / stderr)."""
    code = ("import ctypes, os, sys\n"
            f"lib = ctypes.CDLL({lib_path!r})\n"
            f"f = getattr(lib, {REF_MAIN[name]!r})\n"
            "f.restype = ctypes.c_int\n"
            f"a = [{name!r}.encode()] + [s.encode() for s in {list(args)!r}]\n"
            "argv = (ctypes.c_char_p * (len(a) + 1))(*a, None)\n"
            "rc = f(len(a), argv)\n"
            "os._exit(rc)\n")  # the C++ streams are unbuffered-to-fd or flushed by the reference's own '\n' + exit path
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)


def run_ours(binaries, name, args):
    return subprocess.run([os.path.join(binaries, name), *args], capture_output=True, text=True, timeout=120)


def solver_list_normalised(text):
    return text.replace("Available solvers: ilqr cgd\n", "Available solvers: ilqr\n")


CASES = [
    ("single_track_ocp", ["--help"]),
    ("single_track_ocp", ["-h"]),
    ("single_track_ocp", ["--bogus"]),
    ("single_track_ocp", ["--solver"]),              # missing value
    ("single_track_ocp", ["stray"]),
    ("pendulum_swing_up", ["--help"]),
    ("pendulum_swing_up", ["--agents", "3"]),        # a multi-agent option on a single-OCP program
    ("rocket_max_altitude", ["--help"]),
    ("rocket_max_altitude", ["extra"]),              # this main prints no "Use --help" hint (rocket_max_altitude.cpp:192-196)
    ("multi_agent_single_track", ["--help"]),
    ("multi_agent_single_track", ["--agents", "x"]),
    ("multi_agent_single_track", ["--agents=3", "--agents", "x"]),
    ("multi_agent_single_track", ["--agents", "3x"]),
    ("multi_agent_single_track", ["--max_outer", "q"]),  # underscores in option names are accepted (cli.hpp:15-25)
    ("multi_agent_single_track", ["--max-outer"]),
    ("multi_agent_single_track", ["3", "4"]),        # a second positional
    ("multi_agent_single_track", ["--foo"]),
    ("multi_agent_single_track", ["--agents", "0"]),  # no agents: nothing to solve, a result line with cost 0
    ("multi_agent_lqr", ["--help"]),
    ("multi_agent_lqr", ["--strategy"]),
    ("multi_agent_lqr", ["0"]),
]


@pytest.mark.parametrize("name,args", CASES, ids=[f"{n}:{' '.join(a)}" for n, a in CASES])
def test_command_line_matches_the_reference_main(ref_lib, binaries, name, args):
    ref = run_reference_main(ref_lib, name, args)
    got = run_ours(binaries, name, args)
    assert got.returncode == ref.returncode, (ref.stderr, got.stderr)
    assert got.stderr == ref.stderr
    if args in (["--agents", "0"], ["0"]):  # the result line carries a wall-clock time
        strip = lambda s: " ".join(tok for tok in s.split() if not tok.startswith("time_ms="))  # noqa: E731
        assert strip(got.stdout) == strip(ref.stdout) and "agents=0 cost=0.000000" in got.stdout
    else:
        assert got.stdout == solver_list_normalised(ref.stdout)


# ---- whole runs of the reference's mains: what they print is what the oracle computes ---------------------------------
def parse_output(stdout):
    """First line -> {key: value}; then `<label>_states` / `<label>_controls` blocks (header line, CSV rows, blank line)."""
    import numpy as np

    lines = stdout.splitlines()
    fields = dict(tok.split("=", 1) for tok in lines[0].split())
    blocks, i = {}, 1
    while i < len(lines):
        if lines[i].strip() and "," not in lines[i]:
            label, header, rows = lines[i], lines[i + 1], []
            i += 2
            while i < len(lines) and lines[i].strip():
                rows.append([float(v) for v in lines[i].split(",")])
                i += 1
            blocks[label] = (header, np.array(rows))
        i += 1
    return fields, blocks


def close_at_print_precision(printed, value):
    import numpy as np

    return np.all(np.abs(np.asarray(printed) - np.asarray(value)) <= 0.5e-6 + 1e-12 * np.abs(value))


@pytest.mark.parametrize("name,model,label", [("single_track_ocp", 0, "single_track"), ("pendulum_swing_up", 3, "pendulum"),
                                              ("rocket_max_altitude", 4, "rocket")])
def test_reference_single_ocp_main_prints_the_oracle_result(ref_lib, oracle, name, model, label):
    """`<example> --solver ilqr` of the reference, unmodified and run here, against the oracle with the solver parameters
    of the example's main (libm trig like the reference binary)."""
    from conftest import EXAMPLE_SOLVER_PARAMS

    out = run_reference_main(ref_lib, name, ["--solver", "ilqr"])
    assert out.returncode == 0, out.stderr
    fields, blocks = parse_output(out.stdout)
    x0 = {0: [0.0, 1.0, 0.0, 0.0], 3: [3.141592653589793 - 0.05, 0.0], 4: [0.0, 0.0, 1.0]}[model]
    max_it, tol = EXAMPLE_SOLVER_PARAMS[model]
    import numpy as np

    r = oracle.ilqr_solve_batch(model, np.array([x0]), max_iterations=max_it, tolerance=tol, trig=oracle.TRIG_GLIBC)
    assert fields["solver"] == "ilqr" and close_at_print_precision(float(fields["cost"]), r["cost"][0])
    hdr, X = blocks[label + "_states"]
    n, T, dt = r["X"].shape[2], r["U"].shape[1], oracle.model_dims(model)[3]
    assert hdr == "time," + ",".join(f"x{i}" for i in range(n)) and X.shape == (T + 1, n + 1)
    assert close_at_print_precision(X[:, 0], np.arange(T + 1) * dt) and close_at_print_precision(X[:, 1:], r["X"][0])
    hdr, U = blocks[label + "_controls"]
    assert hdr == "time," + ",".join(f"u{i}" for i in range(r["U"].shape[2])) and close_at_print_precision(U[:, 1:], r["U"][0])


@pytest.mark.parametrize("name,model,strategy,kind", [
    ("multi_agent_single_track", 1, "centralized", 0), ("multi_agent_single_track", 1, "sequential", 1),
    ("multi_agent_single_track", 1, "linesearch", 2), ("multi_agent_single_track", 1, "trustregion", 3),
    ("multi_agent_lqr", 2, "sequential", 1), ("multi_agent_lqr", 2, "centralized", 0)])
def test_reference_multi_agent_main_prints_the_oracle_result(ref_lib, oracle, name, model, strategy, kind):
    """`<example> --agents 3 --strategy S --max-outer 4` of the reference against the oracle's strategy layer: total cost and
    every agent's printed trajectory (blocks in agent-id order)."""
    import numpy as np
    from conftest import circle_x0

    A = 3
    out = run_reference_main(ref_lib, name, ["--agents", str(A), "--solver", "ilqr", "--strategy", strategy, "--max-outer", "4"])
    assert out.returncode == 0, out.stderr
    fields, blocks = parse_output(out.stdout)
    x0 = circle_x0(A)[None] if model == 1 else np.tile([1.0, 0.0, 0.0, 0.0], (1, A, 1))
    r = oracle.strategy_run_batch(kind, model, x0, max_outer=4, max_iterations=100, tolerance=1e-5, trig=oracle.TRIG_GLIBC)
    assert fields["strategy"] == strategy and fields["agents"] == str(A)
    assert close_at_print_precision(float(fields["cost"]), r["total_cost"][0])
    for a in range(A):
        assert close_at_print_precision(blocks[f"agent_{a}_states"][1][:, 1:], r["X"][0, a])
        assert close_at_print_precision(blocks[f"agent_{a}_controls"][1][:, 1:], r["U"][0, a])
