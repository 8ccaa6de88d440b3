#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a markdown table.
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/xyz_launches_summary.md"""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
txt = txt[txt.index('"ID"'):]
agg = {}
n = 0
tot = 0.0
for r in csv.DictReader(io.StringIO(txt)):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
    name = re.sub(r"^ffd::", "", name)
    t = float(r["Metric Value"]) / 1e6
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
    n += 1
    tot += t
print(f"{n} launches captured, {tot:.1f} ms of kernel time.\n")
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{k[:90]}` | {c} | {t:.2f} | {t / c * 1e3:.1f} | {t / tot * 100:.1f}% |")
