"""Per-launch table of the tcgen05 kernels of one training step from `ncu -i report --page raw --csv`:
python scripts/conv_table.py gpurun_out/r02_conv_raw.csv > profiles/r02_conv_tcgen05.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]


def g(r, k, d=0.0):
    try:
        return float(r[idx[k]].replace(",", ""))
    except Exception:
        return d


print("| # | kernel | grid | time us | tensor pipe % | DRAM read MB | DRAM write MB | L2 hit % | L2 throughput % |")
print("|---|---|---|---|---|---|---|---|---|")
tot_t = tot_w = 0.0
for n, r in enumerate(data):
    name = r[idx["Kernel Name"]].split("(")[0].replace("void delora::", "").replace("delora::", "")
    # ncu's csv gives gpu__time_duration in us (unit row) and dram bytes in the unit of rows[1]
    t = g(r, "gpu__time_duration.sum")
    unit_t = rows[1][idx["gpu__time_duration.sum"]]
    if unit_t.startswith("ns"):
        t /= 1000.0
    elif unit_t.startswith("ms"):
        t *= 1000.0
    def mb(k):
        v, u = g(r, k), rows[1][idx[k]]
        return v / 1e6 if u.startswith("byte") else (v / 1e3 if u.startswith("Kbyte") else (v * 1e3 if u.startswith("Gbyte") else v))
    tp = g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    print(f"| {n} | `{name}` | {int(g(r, 'launch__grid_size'))} | {t:.1f} | {tp:.1f} | {mb('dram__bytes_read.sum'):.1f} | "
          f"{mb('dram__bytes_write.sum'):.1f} | {g(r, 'lts__t_sector_hit_rate.pct'):.1f} | "
          f"{g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
    tot_t += t
    tot_w += t * tp
print(f"\nsum {tot_t:.1f} us, time-weighted tensor pipe {tot_w / max(tot_t, 1e-9):.1f} %")
