"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`):
the launches of the last solve in the file, one line each, plus per-kernel totals.  With --json it also writes
per-kernel launches / ms / measured DRAM bytes per launch (bench.py reads that for `roofline.traffic`).

    python profiles/launch_list.py profiles/r01_launches.csv [--json profiles/r01_traffic.json]
"""
import collections
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "second": 1e3}


def short(n):
    m = re.search(r"(forward_coop_kernel|forward_kernel|backward_kernel|prologue_kernel|trial_round_kernel|finish_kernel|aos_to_soa_kernel|"
                  r"soa_to_aos_kernel|rollout_kernel|trust_region_kernel|centralized_kernel)", n)
    s = m.group(0) if m else n[:40]
    t = re.search(r"<([^>]*)>", n)
    return s + ("<" + t.group(1).replace("mas_b200::", "") + ">" if t else "")


def main(path, json_out=None):
    with open(path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"], "ms": 0.0, "dram": 0.0})
        v = float(r["Metric Value"])
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["ms"] = v * TIME[r["Metric Unit"]]
        elif r["Metric Name"].startswith("dram__bytes"):
            d["dram"] += v * SCALE[r["Metric Unit"]]
    launches = list(by.values())
    starts = [i for i, d in enumerate(launches) if "prologue" in d["name"]]
    step = launches[starts[-1]:] if starts else launches
    tot = sum(d["ms"] for d in step)
    agg = collections.OrderedDict()
    for d in step:
        print(f"{short(d['name']):40s} {d['ms']:8.3f} ms  dram {d['dram'] / 1e6:9.1f} MB  grid {d['grid']} block {d['block']}")
        a = agg.setdefault(short(d["name"]).split("<")[0], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["ms"]
        a[2] += d["dram"]
    print(f"total {tot:.3f} ms (serialised, cold cache)")
    for k, (n, ms, b) in agg.items():
        print(f"  {k:20s} launches {n:3d}  {ms:8.3f} ms  {100 * ms / tot:5.1f} %  dram/launch {b / n / 1e6:9.1f} MB")
    if json_out:
        with open(json_out, "w") as f:
            json.dump({k: {"launches": n, "ms": ms, "dram_bytes_per_launch": b / n} for k, (n, ms, b) in agg.items()}, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--json" else None)
