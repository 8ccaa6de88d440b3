#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full) into markdown: key raw metrics + hottest source lines.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.md"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum"]
print(f"Source: `{rep}` (ncu --set full --clock-control none --import-source on, one launch)\n")
print("| metric | unit | value |\n|---|---|---|")
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
for w in want:
    if w in d:
        print(f"| `{w}` | {d[w][0]} | {d[w][1]} |")
for h, (u, v) in d.items():
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            if float(v) >= 0.3:
                print(f"| `{h}` | {u} | {v} |")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = {}; tot = 0; tote = 0
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name", ""): continue
    try: ln = int(r[0]); s = int(r[6]); e = int(r[7])
    except ValueError: continue
    agg[(cur, ln)] = (s, e, r[1].strip()[:110]); tot += s; tote += e
print(f"\nHottest source lines (warp-state samples, {tot} total; executed warp instructions {tote}):\n")
print("| samples | executed | line | source |\n|---|---|---|---|")
for (f, ln), (s, e, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print(f"| {s / tot * 100:.1f}% | {e / tote * 100:.1f}% | {f}:{ln} | `{t.replace('|', '/')}` |")
