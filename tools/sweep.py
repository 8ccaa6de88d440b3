#!/usr/bin/env python
"""BASELINE.json configs 3-4 on N GPUs of one box: python tools/sweep.py --gpus N [--quick]

Runs bench.py --sweep (device-resident samples/s + the final NCCL gather) for
  * NASA-charge (251 x 4), batch-sharded, 16384 samples per GPU                       (config 3)
  * USDroughts (365 x 13) and NASDAQ (252 x 5), GLOBAL batch 2^14 .. 2^20 over the N GPUs, bf16;
    fp32 CUDA-core arithmetic at 2^14 for the tolerance comparison                     (config 4)
  * ECG strong scaling: global 16384 over the N GPUs
and appends one JSON line per configuration to gpurun_out/r2_sweep_N<N>.jsonl.
"""
import argparse, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--quick", action="store_true", help="skip the 2^20 points")
ap.add_argument("--port", type=int, default=29611)
a = ap.parse_args()
N = a.gpus
out = os.path.join(ROOT, "gpurun_out", f"r2_sweep_N{N}.jsonl")
os.makedirs(os.path.dirname(out), exist_ok=True)
configs = [("nasa_charge", ["--batch", "16384"], "fp16", 12)]
for wl in ("droughts", "nasdaq"):
    for e in (14, 16, 18) + (() if a.quick else (20,)):
        g = 1 << e
        if g // N < 1024:
            continue
        steps = 12 if e <= 16 else (6 if e == 18 else 4)
        configs.append((wl, ["--global-batch", str(g), "--sampler-batch", str(1 << 17)], "fp16", steps))
    configs.append((wl, ["--global-batch", str(1 << 14)], "fp32", 4))
configs.append(("ecg", ["--global-batch", "16384"], "fp16", 24))
configs.append(("ecg", ["--batch", "16384"], "fp16", 24))
for wl, extra, prec, steps in configs:
    base = [os.path.join(ROOT, "bench.py"), "--gpus", str(N), "--steps", str(steps), "--warmup", "3", "--sweep",
            "--workload", wl, "--precision", prec] + extra
    cmd = [sys.executable] + base if N == 1 else [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={N}", "--master-addr", "127.0.0.1",
        "--master-port", str(a.port)] + base
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = None
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            line = json.loads(ln)
    rec = {"workload": wl, "args": extra, "precision": prec, "n_gpus": N, "wall_s": round(time.time() - t0, 1),
           "result": line, "rc": r.returncode}
    if line is None:
        rec["stderr_tail"] = r.stderr[-600:]
    with open(out, "a") as f:
        f.write(json.dumps(rec) + "\n")
    v = line and (round(line["value"], 1), round(line["ms_per_step"], 2), line["config"].get("final_gather_ms"))
    print(wl, extra, prec, "->", v, f"({rec['wall_s']} s)", flush=True)
