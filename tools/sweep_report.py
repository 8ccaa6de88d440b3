#!/usr/bin/env python
"""Merge gpurun_out/r2_sweep_N*.jsonl (tools/sweep.py) into profiles/r2_sweep.json + a markdown table on stdout."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r2_sweep_N*.jsonl"))):
    for ln in open(f):
        r = json.loads(ln)
        res = r.get("result")
        if not res:
            continue
        c = res["config"]
        rows.append({"workload": r["workload"], "precision": r["precision"], "n_gpus": r["n_gpus"],
                     "scaling": res["scaling"], "global_batch": c["global_batch"], "batch_per_gpu": c["batch_per_gpu"],
                     "sampler_batch": c["sampler_batch"], "steps_timed": res["steps"],
                     "samples_per_s": round(res["value"], 1), "ms_per_step": round(res["ms_per_step"], 3),
                     "ms_per_pure_cache_step": round(c["ms_per_step_after_step0"], 3),
                     "final_gather_ms": None if c["final_gather_ms"] is None else round(c["final_gather_ms"], 3),
                     "step_frac_of_sustained_bf16_peak": None if c["step_frac_of_sustained_bf16_peak"] is None
                     else round(c["step_frac_of_sustained_bf16_peak"], 3),
                     "sm_mhz": res["clocks"].get("sm_mhz"), "reasons": res["clocks"].get("reasons")})
base = {(r["workload"], r["precision"], r["global_batch"] // r["n_gpus"] if r["scaling"] == "weak" else r["global_batch"],
         r["scaling"]): r["samples_per_s"] for r in rows if r["n_gpus"] == 1}
for r in rows:
    k1 = (r["workload"], r["precision"], r["global_batch"], "strong")
    kw = (r["workload"], r["precision"], r["batch_per_gpu"], "weak")
    ref = base.get(k1) if r["scaling"] == "strong" else base.get(kw)
    r["speedup_vs_1gpu"] = round(r["samples_per_s"] / ref, 2) if ref else None
out = {"what": "BASELINE.json configs 3-4 (+ ECG strong/weak): device-resident samples/s for a 1000-step run = global batch / "
               "(1000 x mean step time over the timed schedule steps 0..K-1, step 0 = full recompute), one 8xB200 box; "
               "final_gather_ms = ONE all_gather_into_tensor of every rank's last sampler batch into a preallocated, "
               "once-warmed buffer; tools/sweep.py -> bench.py --sweep; strong = --global-batch split over the GPUs",
       "rows": rows}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_sweep.json"), "w"), indent=1)
print("| workload | prec | GPUs | global batch | samples/s | x 1 GPU | ms/step | gather ms | step / sustained peak |\n|---|---|---|---|---|---|---|---|---|")
for r in sorted(rows, key=lambda r: (r["workload"], r["precision"], r["global_batch"], r["n_gpus"], r["scaling"])):
    print(f"| {r['workload']} | {r['precision']} | {r['n_gpus']} ({r['scaling']}) | {r['global_batch']} | {r['samples_per_s']} | "
          f"{r['speedup_vs_1gpu']} | {r['ms_per_step']} | {r['final_gather_ms']} | {r['step_frac_of_sustained_bf16_peak']} |")
