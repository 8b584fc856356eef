"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_list.py gpurun_out/launches.csv [skip_first_n]"""
import csv, sys, collections, re
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v_us, r.get("Grid Size", "")))
rows = rows[int(sys.argv[2]) if len(sys.argv) > 2 else 0:]
agg = collections.OrderedDict()
for n, v, g in rows:
    a = agg.setdefault((n, g), [0, 0.0, 1e30, 0.0]); a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(v for _, v, _ in rows)
print(f"{len(rows)} launches, {tot:.1f} us GPU time")
for (n, g), (c, s, lo, hi) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*s/tot:5.1f}%  {c:5d} x  avg {s/c:8.2f} us  (min {lo:7.2f}, max {hi:7.2f})  grid {g:>14s}  {n[:60]}")
