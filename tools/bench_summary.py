"""One-line digest of bench.py JSON lines: python tools/bench_summary.py profiles/r02_bench_*gpu.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        s = open(f).read()
        d = json.loads(s[s.index('{"metric'):].strip().splitlines()[-1])
        eng = d.get("engine", d.get("config", {}))
        par = d.get("parity") or {}
        print(f, f"N={d['n_gpus']} {d['scaling']}", f"value {d['value'] / 1e6:.2f} M/s", f"e2e {d['e2e']['value'] / 1e6:.2f} M/s",
              f"one solve {eng.get('single_solve_ms', float('nan')):.2f} ms", f"depth {eng.get('solves_in_flight')}",
              f"parity {par.get('bit_equal_problems')}/{par.get('problems_checked')}",
              f"weak {d['weak']['value'] / 1e6:.2f} M/s" if d.get("weak") else "")
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
