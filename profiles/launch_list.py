"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the launches of the last solve
in the file, one line each, plus per-kernel totals.  usage: python profiles/launch_list.py <csv>"""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r"(forward_kernel|backward_kernel|prologue_kernel|aos_to_soa_kernel|soa_to_aos_kernel|rollout_kernel|trust_region_kernel)", n)
    s = m.group(0) if m else n[:40]
    t = re.search(r"<([^>]*)>", n)
    return s + ("<" + t.group(1).replace("mas_b200::", "") + ">" if t else "")


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    seq = [(short(r["Kernel Name"]), float(r["Metric Value"]) / 1e6, r["Grid Size"], r["Block Size"]) for r in rows]
    starts = [i for i, s in enumerate(seq) if s[0].startswith("prologue")]
    start = starts[-1] if starts else 0
    tot = 0.0
    agg = collections.defaultdict(float)
    for name, ms, grid, block in seq[start:]:
        print(f"{name:42s} {ms:8.3f} ms  grid {grid} block {block}")
        tot += ms
        agg[name.split("<")[0]] += ms
    print(f"total {tot:.3f} ms (serialised, cold cache)")
    for k, v in agg.items():
        print(f"  {k:20s} {v:8.3f} ms  {100 * v / tot:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
