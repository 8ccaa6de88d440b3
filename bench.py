#!/usr/bin/env python
"""bench.py -- samples/sec of ECG frequency-domain reverse-SDE sampling with E2-CRF caching.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): ECG (L=187, C=1), Hydra-default transformer score model
(d=72, 12 heads, 10 layers, FFN 2048, random init under seed 42), VP-SDE Euler-Maruyama on the
1000-step time grid, E2-CRF cache with benchmark_cache.py's default kwargs (K=5, R=10), batch 16384
per GPU, synthetic prior, Philox noise drawn in-kernel.

A bench "step" is ONE reverse-diffusion step (score-model forward + SDE update) over the whole
batch; the timed region runs schedule steps 0..K-1 (step 0 is the full-recompute step, step 500
the refresh step, the rest pure-cache steps).  value = samples/sec for a 1000-step run
= batch / (1000 * mean step time); with K = 1000 that is the exact wall time of a full run.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_DIFFUSION_STEPS = 1000
ECG = dict(max_len=187, n_channels=1, d_model=72, num_layers=10, n_head=12)
D_FF = 2048
METRIC = "samples/sec (ECG freq-domain, 1000 steps, E2-CRF)"
# (max_len, n_channels) of the reference's datamodules (src/fdiff/dataloaders/datamodules.py:194-203 ECG,
# :469-476 NASA-charge, :406-410 NASDAQ, :530-538 USDroughts); the score model is the Hydra default for all of them
SHAPES = {"ecg": ("ECG", 187, 1), "nasa_charge": ("NASA-charge", 251, 4), "nasdaq": ("NASDAQ", 252, 5),
          "droughts": ("USDroughts", 365, 13)}


def set_workload(name: str) -> None:
    """BASELINE.json configs[1] is ECG (the default and the driver's bench line); configs[2-3] name the other shapes."""
    global METRIC, WORKLOAD
    label, L, C = SHAPES[name]
    ECG.update(max_len=L, n_channels=C)
    METRIC = f"samples/sec ({label} freq-domain, 1000 steps, E2-CRF)"
    WORKLOAD = (f"{label} L={L} C={C} freq-domain, d=72 H=12 NL=10 F=2048 random init (seed 42), VP-SDE 1000-step grid, "
                "E2-CRF K=5 R=10")


# the FFN + LayerNorm kernel the library launches for the default model: CTA pairs unless FFD_FFN_CTA2=0
FFN_KERNEL_NAME = ("ffn_ln_tc2_kernel<80, 72> (cta_group::2 pairs)" if os.environ.get("FFD_FFN_CTA2", "1")[:1] == "1"
                   else "ffn_ln_tc_kernel<80>")


WORKLOAD = ("ECG L=187 C=1 freq-domain, d=72 H=12 NL=10 F=2048 random init (seed 42), VP-SDE 1000-step grid, "
            "E2-CRF K=5 R=10")
_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that lands on fd 1 (NCCL prints its version banner there,
    torch.distributed.run its warnings) is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


def flops_per_sample_step(n_own: int, L=None, C=None, d=72, F=D_FF, NL=10) -> float:
    """NL*L*(8d^2 + 4Ld + 4dF) + 4LCd for a full step; the pure-cache step skips K/V (4d^2);
    a mixed step recomputes K/V for n_own rows."""
    L = ECG["max_len"] if L is None else L
    C = ECG["n_channels"] if C is None else C
    if n_own < 0 or n_own == L or n_own > 0.8 * L:
        return NL * L * (8 * d * d + 4 * L * d + 4 * d * F) + 4 * L * C * d
    return NL * (L * (4 * d * d + 4 * L * d + 4 * d * F) + 4 * n_own * d * d) + 4 * L * C * d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.source = "nvml"
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _nvml_handle(self):
        """In-process NVML (no nvidia-smi process per sample: spawning one initialises every GPU of the box and was
        measurably perturbing multi-GPU runs).  The device is matched by UUID, falling back to the index."""
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None, None

    def _sample_nvml(self, nv, h):
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def _run(self):
        nv, h = self._nvml_handle()
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.rows.append(self._sample_nvml(nv, h))
                else:
                    self.source = "nvidia-smi"
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) == 6:
                        self.rows.append(parts)
            except Exception:
                nv = None  # fall back to nvidia-smi
                self.source = "nvidia-smi"
            self._stop.wait(0.1 if nv is not None else 0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def build_model(precision: str):
    import torch

    import fastfourierdiffusion_b200 as F

    torch.manual_seed(42)
    sched = F.VPScheduler(fourier_noise_scaling=True)
    model = F.ScoreModule(noise_scheduler=sched, fourier_noise_scaling=True, precision=precision, **ECG)
    sched.set_noise_scaling(ECG["max_len"])
    return model


# ------------------------------------------------------------------------------------------------
def cpu_backend():
    """"reference": the unmodified reference classes from oracle/_ref (vendored by build() where the upstream checkout
    exists; travels to the GPU box with the snapshot); "port": oracle/torch_port.py when that copy is absent."""
    try:
        from oracle import ref_runner

        if ref_runner.available():
            return "reference"
    except Exception:
        pass
    return "port"


def cpu_steps(kind, state_dict, batch, steps, n_threads):
    """Seconds per schedule step (dict step -> s) of the CPU implementation on a fresh batch and a fresh cache."""
    import torch

    torch.set_num_threads(n_threads)
    if kind == "reference":
        from oracle import ref_runner

        t, _ = ref_runner.time_reference_steps(dict(ECG), batch, steps, N_DIFFUSION_STEPS, n_threads=n_threads)
        return t
    from oracle.torch_port import TorchPort, time_cached_steps

    port = TorchPort(state_dict, ECG["max_len"], ECG["n_channels"], ECG["n_head"], "vp", True)
    t, _ = time_cached_steps(port, batch, steps, N_DIFFUSION_STEPS)
    return t


def cpu_port_throughput(state_dict, budget_s: float, n_threads: int, steps=None, warmup=0):
    """samples/sec of the reference's CPU implementation (oracle/_ref, else the torch-CPU port) on a bounded sample;
    returns (value, sample description, ms/step, kind)."""
    kind = cpu_backend()
    # calibrate the batch so the run fits the budget
    probe = cpu_steps(kind, state_dict, 4, [0, 1], n_threads)
    per_sample = max(probe[1], 1e-4) / 4
    if steps is None:  # cpu_baseline leg: full + 3 pure + refresh step, extrapolated over the schedule
        b = int(max(4, min(64, budget_s / (6 * per_sample))))
        t = cpu_steps(kind, state_dict, b, [0, 1, 2, 3, 500], n_threads)
        pure = (t[1] + t[2] + t[3]) / 3
        total = t[0] + t[500] + (N_DIFFUSION_STEPS - 2) * pure
        return b / total, (f"batch {b}: schedule steps 0 (full), 1-3 (pure cache), 500 (refresh) timed, "
                           f"1000-step run = full + refresh + 998 x mean pure step"), total / N_DIFFUSION_STEPS * 1e3, kind
    k = len(steps)
    b = int(max(1, min(64, budget_s / ((k + warmup + 2) * per_sample))))
    t = cpu_steps(kind, state_dict, b, list(steps), n_threads)  # the probe above was the warm-up (thread pool, allocator)
    mean = sum(t.values()) / k
    return b / (mean * N_DIFFUSION_STEPS), f"batch {b}: schedule steps 0..{k - 1} of the 1000-step grid", mean * 1e3, kind


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The upstream package is pure
    Python and absent on the GPU box, so the torch-CPU port of oracle/ stands in (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    model = build_model("fp32")
    sd = {k: v for k, v in model.state_dict().items()}
    value, sample, ms, kind = cpu_port_throughput(sd, budget_s=150.0, n_threads=cores, steps=list(range(args.steps)),
                                                  warmup=args.warmup)
    # informational: the same ATen op sequence (stock torch eager, TF32 matmuls) on one GPU of the box -- what the
    # unmodified reference's operators would reach on this hardware.  Not the arm's value.
    on_gpu = None
    try:
        if torch.cuda.is_available():
            from oracle.torch_port import TorchPort, time_cached_steps

            torch.set_float32_matmul_precision("high")  # TF32 matmuls, as the reference sets them (cmd/sample.py:24)
            gp = TorchPort(sd, ECG["max_len"], ECG["n_channels"], ECG["n_head"], device="cuda")
            on_gpu = {"unit": "samples/s",
                      "what": "oracle/torch_port.py (the reference's ATen op sequence) in torch eager on cuda:0, TF32 "
                              "matmuls as in cmd/sample.py:24", "batches": {}}
            for gb in (4096, args.batch):  # the bench batch too, when the eager path fits it in memory
                try:
                    gp.k_cache = gp.v_cache = None
                    time_cached_steps(gp, gb, [0, 1], N_DIFFUSION_STEPS)  # warm-up (cuBLAS handles, allocator)
                    gp.k_cache = gp.v_cache = None
                    tt, _ = time_cached_steps(gp, gb, [0, 1, 2, 3, 4, 5], N_DIFFUSION_STEPS)
                    pure = sum(tt[i] for i in (1, 2, 3, 4, 5)) / 5
                    on_gpu["batches"][str(gb)] = {"value": gb / (pure * N_DIFFUSION_STEPS),
                                                  "ms_per_pure_cache_step": pure * 1e3, "ms_full_step": tt[0] * 1e3}
                except Exception as exc:  # out of memory at the large batch: recorded, not fatal
                    on_gpu["batches"][str(gb)] = {"error": str(exc)[:120]}
                    torch.cuda.empty_cache()
            best = [v["value"] for v in on_gpu["batches"].values() if "value" in v]
            on_gpu["value"] = max(best) if best else None
            del gp
            torch.cuda.empty_cache()
    except Exception as exc:
        on_gpu = {"error": str(exc)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "arm": ("the UNMODIFIED reference classes (oracle/_ref, vendored from the upstream checkout by build()) "
                           "driven step by step as DiffusionSampler.sample does, CPU fp32" if kind == "reference" else
                           "CPU torch port of the reference path (oracle/torch_port.py), fp32"),
                   "torch_threads": torch.get_num_threads(), "reference_ops_on_gpu": on_gpu},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        entry.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import fastfourierdiffusion_b200 as F
    from fastfourierdiffusion_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda", local)
    if world > 1:
        from fastfourierdiffusion_b200.distributed import warm_communicator

        warm_communicator(dev)  # NCCL connects lazily: keep the set-up out of every timed collective
    strong = args.global_batch > 0
    B = args.global_batch // world if strong else args.batch
    SB = min(B, args.sampler_batch) if args.sampler_batch > 0 else B   # samples per sampler batch (one workspace)
    n_chunks = (B + SB - 1) // SB
    K, W = args.steps, args.warmup
    model = build_model(args.precision).to(dev).eval()
    cpu_state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sampler = F.DiffusionSampler(model, SB, use_cache=True, cache_kwargs={}, noise="philox", seed=42 + rank)
    sched = model.noise_scheduler
    sched.set_timesteps(N_DIFFUSION_STEPS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident run -------------------------------------------------------------------
    # B samples per GPU in n_chunks sampler batches of SB (the reference's sample_batch_size loop, sampler.py:157):
    # every sampler batch runs schedule steps 0..K-1
    sizes = [min(SB, B - c * SB) for c in range(n_chunks)]
    X0s = [sampler.sample_prior(n, c) for c, n in enumerate(sizes)]
    model.cache.reset()
    sampler.run_steps(X0s[0], 0, step_begin=0, step_end=max(W, 1))  # warm-up: full step + pure steps
    model.cache.reset()
    barrier()
    n0 = lib.ffd_launch_count()
    ev0, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_chunks)]
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for c in range(n_chunks):
            model.cache.reset()
            evm[c][0].record()
            X = sampler.run_steps(X0s[c], 0, batch_idx=c, step_begin=0, step_end=1)  # step 0: full recompute + cache fill
            evm[c][1].record()
            if K > 1:
                X = sampler.run_steps(X, 0, batch_idx=c, step_begin=1, step_end=K)
        ev1.record()
        barrier()
    launches = int(lib.ffd_launch_count() - n0)
    ms = torch.tensor([ev0.elapsed_time(ev1), sum(a.elapsed_time(b) for a, b in evm)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total, ms_step0 = float(ms[0].item()), float(ms[1].item())
    ms_rest = (ms_total - ms_step0) / max(K - 1, 1)
    assert torch.isfinite(X).all()
    ms_per_step = ms_total / K
    value = world * B / (ms_per_step * 1e-3 * N_DIFFUSION_STEPS)
    del X0s

    # ---- end to end through the public API with host buffers ------------------------------------
    # DiffusionSampler.sample(): prior noise drawn on the host, copied to the device, K steps with
    # per-step torch noise, final samples copied back to the host (sampler.py:105-215).
    # N > 1: every rank samples its shard (sample_on_device), ONE NCCL all-gather assembles [world * B, L, C] on every
    # rank (the path's only collective) and rank 0 reads the result to the host -- all inside the timed region.
    from fastfourierdiffusion_b200.distributed import gather_samples

    e2e_sampler = F.DiffusionSampler(model, SB, use_cache=True, cache_kwargs={}, noise="torch")
    Ke = min(K, args.e2e_steps) if args.e2e_steps > 0 else K
    gather_ms = None
    gather_skew_ms = None
    gbuf = None
    if world > 1 and not args.sweep:
        # the gathered result lives in a buffer allocated BEFORE the timed region (a first-touch cudaMalloc of
        # world x 12 MB inside it was most of the "53.6 ms gather" of the round-1 scaling run), and one untimed gather
        # of the same size warms NCCL's channels for that message size
        gbuf = torch.empty((world * B, ECG["max_len"], ECG["n_channels"]), device=dev, dtype=torch.float32)
        gather_samples_warm = torch.zeros((B, ECG["max_len"], ECG["n_channels"]), device=dev)
        from fastfourierdiffusion_b200.distributed import gather_samples as _gs

        _gs(gather_samples_warm, world * B, out=gbuf)
        torch.cuda.synchronize()
        del gather_samples_warm
        # rank 0 reads the gathered result into pinned host memory (allocated outside the timed region)
        host_out = torch.empty(gbuf.shape, dtype=torch.float32, pin_memory=True) if rank == 0 else None
    if args.sweep:
        return finish_sweep_line(args, world, rank, B, SB, n_chunks, K, W, value, ms_per_step, ms_step0, ms_rest, launches,
                                 clocks, strong, X, dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        out = e2e_sampler.sample(B, num_diffusion_steps=Ke)
        d2h = B
    else:
        Xd = e2e_sampler.sample_on_device(B, num_diffusion_steps=Ke)
        g0.record()
        allX = gather_samples(Xd, world * B, out=gbuf)
        g1.record()
        out = host_out.copy_(allX) if rank == 0 else allX[:0].cpu()
        d2h = world * B if rank == 0 else 0
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        gm = torch.tensor([g0.elapsed_time(g1)], device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_skew_ms = float(gm.item())  # inside the e2e region: the collective PLUS the wait for the slowest rank
        assert rank != 0 or out.shape[0] == world * B
        # the collective itself: the same gather again with every rank entering together (outside the e2e region)
        barrier()
        g0.record()
        allX = gather_samples(Xd, world * B, out=gbuf)
        g1.record()
        torch.cuda.synchronize()
        gm = torch.tensor([g0.elapsed_time(g1)], device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_ms = float(gm.item())
        del allX, Xd, gbuf
    assert tuple(out.shape[1:]) == (ECG["max_len"], ECG["n_channels"]) and out.device.type == "cpu"
    e2e_value = world * B / (float(t_e2e.item()) / Ke * N_DIFFUSION_STEPS)
    bytes_io = B * ECG["max_len"] * ECG["n_channels"] * 4
    d2h_bytes = d2h * ECG["max_len"] * ECG["n_channels"] * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused FFN + LN), timed alone with CUDA events -------------
    pk = peaks()
    Bk = min(B, 16384)   # the kernel-level timings below use one bench-sized batch of rows
    M = Bk * ECG["max_len"]
    arena, packed = model.device_weights()
    xbuf = torch.randn(M, ECG["d_model"], device=dev)
    ybuf = torch.empty_like(xbuf)
    scratch = torch.empty_like(xbuf) if args.precision == "fp32" else None
    prec = 1 if args.precision == "fp16" else 0
    st = _lib.stream_handle(dev)

    def ffn():
        _lib.check(lib.ffd_ffn_ln(model.dims, prec, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(xbuf),
                                  _lib.ptr(ybuf), _lib.ptr(scratch), M, st))

    for _ in range(3):
        ffn()
    torch.cuda.synchronize()
    reps = 10
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(reps):
        ffn()
    r1.record()
    torch.cuda.synchronize()
    ffn_ms = r0.elapsed_time(r1) / reps
    ffn_flops = 4.0 * M * ECG["d_model"] * D_FF
    achieved = ffn_flops / (ffn_ms * 1e-3) / 1e12
    # second kernel of a pure-cache layer: the attention block.  It is bound by MUFU.EX2 (one exponential per
    # (token, head, key): 16 lanes/clk/SM), not by HBM or the tensor pipe, so it gets its own informational entry.
    attn_entry = None
    if args.precision == "fp16":
        kc, vc = model.cache._k_store, model.cache._v_store
        wsb = model.workspace(Bk)

        def attn():
            _lib.check(lib.ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 3, _lib.ptr(kc),
                                              _lib.ptr(vc), _lib.ptr(xbuf), _lib.ptr(ybuf), _lib.ptr(wsb), wsb.numel(),
                                              M, st))

        for _ in range(3):
            attn()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            attn()
        a1.record()
        torch.cuda.synchronize()
        attn_ms = a0.elapsed_time(a1) / reps
        n_exp = float(M) * ECG["n_head"] * ECG["max_len"]
        mufu_peak = 148 * 16 * (clocks.summary().get("sm_mhz") or 1965.0) * 1e6
        attn_flops = M * (4.0 * 72 * 72 + 4.0 * ECG["max_len"] * 72)
        attn_entry = {"kernel": "attn_cached_tc_kernel<80, 3> (+ K/V image pack)", "bound": "sfu (MUFU.EX2)",
                      "achieved": n_exp / (attn_ms * 1e-3) / 1e12, "peak": mufu_peak / 1e12, "unit": "Texp/s",
                      "frac": n_exp / (attn_ms * 1e-3) / mufu_peak, "ms_per_launch": attn_ms,
                      "tensor_tflops_algorithmic": attn_flops / (attn_ms * 1e-3) / 1e12,
                      "peak_source": "148 SMs x 16 MUFU lanes x SM clock under load"}
    peak = pk["bf16"] if args.precision == "fp16" else 72.0  # fp32 FMA peak: 148 SMs x 128 lanes x 2 x 1.9 GHz
    # DRAM traffic per launch comes from an ncu --set full capture (it cannot be measured inside a timed run): the
    # committed summary of that capture, valid for the ECG M it was taken at; labelled as static
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "ffn_traffic.json")
    if os.path.exists(tpath) and M == 16384 * 187:
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            traffic_source = "profiles/ffn_traffic.json (static: ncu --set full capture of this kernel at this M)"
        except Exception:
            traffic = None
    # whole-step figure: algorithmic FLOPs of the timed steps over the measured time
    n_own = [int(v) for v in sampler.step_tables(0)[1]][:K]
    step_flops = sum(flops_per_sample_step(n) for n in n_own) * B
    step_tflops = step_flops / (ms_total * 1e-3) / 1e12

    # ---- informational: the device-decided event schedule (paper Algorithm 1) on the same batch ---------------
    event_info = None
    try:
        ev_sampler = F.DiffusionSampler(model, SB, use_cache=True, noise="philox", seed=42 + rank,
                                        cache_kwargs=dict(K=5, R=50, tau_0=0.3, tau_warn=0.5, event_driven=True,
                                                          event_probe_ratio=0.02))
        model.cache.reset()
        Xe = ev_sampler.run_steps(ev_sampler.sample_prior(SB), 0, step_begin=0, step_end=3)
        torch.cuda.synchronize()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_ev = min(20, max(K - 3, 1))
        q0.record()
        Xe = ev_sampler.run_steps(Xe, 0, step_begin=3, step_end=3 + n_ev)
        q1.record()
        torch.cuda.synchronize()
        est = model.cache.get_event_stats()
        event_info = {"ms_per_step": q0.elapsed_time(q1) / n_ev, "steps": n_ev,
                      "recompute_fraction": est.get("recompute_fraction"),
                      "kwargs": "K=5 R=50 tau_0=0.3 tau_warn=0.5 probe=0.02; decisions on the device, no host sync"}
    except Exception as exc:  # informational only
        event_info = {"error": str(exc)[:200]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample, _, kind = cpu_port_throughput(cpu_state, budget_s=20.0, n_threads=cores)
        cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None,
        "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "noise": "Philox in-kernel",
                   "batch_per_gpu": B, "global_batch": world * B, "sampler_batch": SB,
                   "sampler_batches_per_gpu": n_chunks, "schedule_steps_timed": f"0..{K - 1}",
                   "value_definition": "global_batch / (1000 x mean step time over the timed steps)",
                   "l2": "per-step activations [B*L, 72] fp32 = %.0f MB >> 126 MB L2 (inputs larger than L2)" % (
                       M * 72 * 4 / 1e6),
                   "parallelism": f"batch-sharded x{world}, no per-step collective",
                   "final_gather_ms": gather_ms, "final_gather_incl_rank_skew_ms": gather_skew_ms,
                   "ms_step0_full_recompute": ms_step0,
                   "ms_per_step_after_step0": ms_rest, "step_tflops_algorithmic": step_tflops,
                   "step_frac_of_sustained_bf16_peak": step_tflops / pk["bf16_sustained"],
                   "event_driven_schedule": event_info},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": bytes_io / Ke,
                "d2h_bytes_per_step": d2h_bytes / Ke, "steps": Ke,
                "api": ("DiffusionSampler.sample(): host prior noise -> H2D -> steps (torch device noise) -> D2H"
                        if world == 1 else
                        "per rank DiffusionSampler.sample_on_device() (host prior noise -> H2D -> steps) -> ONE NCCL "
                        "all_gather_into_tensor of the samples -> D2H of [world*B, L, C] on rank 0, all timed")},
        "gpu_launches": launches,
        "roofline": {"kernel": FFN_KERNEL_NAME if args.precision == "fp16" else "ffn_kernel+add_ln_kernel",
                     "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "peak_source": pk["src"] + " (burst)",
                     "flops_per_launch": ffn_flops, "ms_per_launch": ffn_ms,
                     "how": f"{reps} back-to-back launches over M={M} rows between CUDA events on the launch stream"},
        "roofline_other": [attn_entry] if attn_entry else [],
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def finish_sweep_line(args, world, rank, B, SB, n_chunks, K, W, value, ms_per_step, ms_step0, ms_rest, launches, clocks,
                      strong, X, dev):
    """--sweep: the device-resident throughput and (N > 1) the final NCCL gather only -- one short JSON line per
    configuration for tools/sweep.py (profiles/r2_sweep.json).  No e2e / roofline / CPU legs: those belong to the
    driver's bench line of the default workload."""
    import torch
    import torch.distributed as dist

    gather_ms = None
    if world > 1:
        from fastfourierdiffusion_b200.distributed import gather_samples

        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gbuf = torch.empty((world * X.shape[0],) + tuple(X.shape[1:]), device=dev, dtype=torch.float32)
        gather_samples(X, world * X.shape[0], out=gbuf)  # untimed: allocation and NCCL set-up for this message size
        dist.barrier()
        torch.cuda.synchronize()
        g0.record()
        allX = gather_samples(X, world * X.shape[0], out=gbuf)
        g1.record()
        torch.cuda.synchronize()
        gm = torch.tensor([g0.elapsed_time(g1)], device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_ms = float(gm.item())
        assert allX.shape[0] == world * X.shape[0]
        del allX
    if rank == 0:
        flops = sum(flops_per_sample_step(0 if i else ECG["max_len"]) for i in range(K)) * B
        emit({"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
              "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
              "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
              "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "sampler_batch": SB,
                         "sampler_batches_per_gpu": n_chunks, "ms_step0_full_recompute": ms_step0,
                         "ms_per_step_after_step0": ms_rest, "final_gather_ms": gather_ms,
                         "gathered_samples_per_rank": int(X.shape[0]),
                         "step_frac_of_sustained_bf16_peak": (flops / (ms_per_step * K * 1e-3) / 1e12
                                                              / peaks()["bf16_sustained"]) if args.precision == "fp16" else None},
              "clocks": clocks.summary(), "gpu_launches": launches})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="samples per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="> 0: total samples over all GPUs (strong scaling: --batch is ignored)")
    ap.add_argument("--sampler-batch", type=int, default=0,
                    help="> 0: samples per sampler batch (sample_batch_size); a GPU's batch runs in chunks of this size")
    ap.add_argument("--workload", default="ecg", choices=sorted(SHAPES))
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16: tcgen05 GEMMs on fp16 operands, fp32 accumulate (bf16 = its former name); fp32: CUDA cores")
    ap.add_argument("--e2e-steps", type=int, default=0, help="cap the steps of the e2e leg (0 = same as --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="short line: device-resident throughput (+ final gather) only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    set_workload(args.workload)
    if args.precision == "bf16":
        args.precision = "fp16"
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
