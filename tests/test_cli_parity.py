"""The example programs' command line against the reference's OWN main()s, on command lines that need no device.

oracle/_ref/libref.so holds the reference's five example sources unmodified, their `main` renamed by the preprocessor
(oracle/ref/ex_*.cpp), so the reference's argument parser (examples/cli.hpp), usage text and error messages can be run
here and compared byte for byte with examples/bin/* (CLI + result-line format).  The one intended
difference is the solver list: `cgd` is not on the device path.  The result line and the CSV blocks of a real solve
are compared on the GPU (tests/test_gpu_facade.py).
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
