"""Top source lines of an ncu capture taken with --import-source on.

  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > page.csv
  python profiles/source_hotspots.py page.csv [N]
"""
import collections
import csv
import os
import sys

def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data = collections.defaultdict(list)
path = kern = hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        path = os.path.basename(r[1])
    elif len(r) >= 2 and r[0] == "Function Name":
        kern = r[1].split("(")[0].replace("void mas_b200::", "")
    elif r and r[0] == "Line No":
        hdr = r
    elif kern and r and r[0].strip().isdigit():
        data[kern].append((path, r))
for k, v in data.items():
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_s = sum(num(r[i_s]) for _, r in v)
    tot_i = sum(num(r[i_i]) for _, r in v)
    print(f"===== {k}: {tot_i} warp instructions, {tot_s} stall samples")
    per_file = collections.Counter()
    for p, r in v:
        per_file[p] += num(r[i_i])
    print("  per file:", {p: f"{100 * n / tot_i:.1f}%" for p, n in per_file.most_common()})
    for p, r in sorted(v, key=lambda pr: -num(pr[1][i_i]))[:top]:
        print(f"  {p:>20}:{r[0]:<5} inst {100 * num(r[i_i]) / tot_i:5.1f}%  samples {100 * num(r[i_s]) / tot_s:5.1f}%  {r[1].strip()[:100]}")
