"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:  python tools/launch_summary.py launches.csv"""
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = {}
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v_us
tot = sum(a[1] for a in agg.values())
print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms in total (serialised, cold cache)\n")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {cnt} | {us / 1e3:.3f} | {us / cnt:.1f} | {us / tot * 100:.1f}% |")
