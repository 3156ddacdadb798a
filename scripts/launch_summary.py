"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,sm__pipe_tensor_cycles_active...] --csv):
per-kernel totals of the LAST step (the list must hold `steps` identical steps).  python scripts/launch_summary.py file.csv [steps]"""
import collections
import csv
import sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
idx = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in data:
    if len(r) < len(hdr):
        continue
    d = launch.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]]})
    d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
ids = sorted(launch)
marks = [i for i in ids if "project_scatter" in launch[i]["name"]]
if marks:                       # one step = from the last projection launch to the end of the list
    last = [i for i in ids if i >= marks[-1]]
    n = len(last)
else:
    n = len(ids) // steps
    last = ids[-n:]
tot = collections.OrderedDict()
for i in last:
    d = launch[i]
    nm = d["name"].split("(")[0][-56:]
    t = tot.setdefault(nm, [0, 0.0, 0.0])
    us = d.get("gpu__time_duration.sum", 0) / 1000.0
    t[0] += 1
    t[1] += us
    t[2] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * us
s = sum(v[1] for v in tot.values())
print(f"launches per step {n}, sum {s:.1f} us (serialised, cold caches: compare SHARES)")
print("| kernel | launches | total us | share | tensor pipe % (time-weighted) |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / s:.1f} % | {v[2] / max(v[1], 1e-9):.1f} |")
