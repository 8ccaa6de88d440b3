"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes.

TEST INFRASTRUCTURE ONLY (see oracle/fdiff_oracle.py header).  Run in the build
container, where the upstream checkout is mounted:

    python oracle/make_golden.py            # writes tests/golden/*.npz

The GPU box has no upstream checkout, so the fixtures are committed.  Two model
sizes are used:

* ``tiny``  (L=16, C=2, d=12, H=2, NL=2): weights are stored in the fixture.
* ``ecg``   (L=187, C=1, d=72, H=12, NL=10 -- the Hydra default): 12.8 MB of
  weights are NOT stored; they are re-derived from ``torch.manual_seed(seed)``
  through the same constructor order plus ``ref_harness.redraw_layers`` and the
  fixture keeps a per-tensor checksum so a drift is detected, not silently used.

Each trace holds the reference's own noise draws (prior + per-step) so the
oracle and the CUDA path can be driven with injected noise.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def checksums(sd):
    keys = sorted(sd)
    return np.array(keys), np.array([[float(np.abs(sd[k].astype(np.float64)).sum()),
                                      float((sd[k].astype(np.float64) ** 2).sum())] for k in keys])


def run_trace(ref, model, *, batch, n_samples, n_steps, use_cache, cache_kwargs, keep_steps):
    """Drive DiffusionSampler.sample and record noise, states, scores."""
    sampler = ref.DiffusionSampler(model, sample_batch_size=batch, use_cache=use_cache,
                                   cache_kwargs=cache_kwargs)
    zlog, slog, sclog = [], [], []
    orig_fwd = model.forward

    def fwd(*a, **k):
        out = orig_fwd(*a, **k)
        sclog.append((out[0] if isinstance(out, tuple) else out).detach().clone())
        return out

    model.forward = fwd
    try:
        with rh.record_randn(zlog), rh.record_steps(sampler, slog):
            X = sampler.sample(n_samples, n_steps)
    finally:
        model.forward = orig_fwd
    n_batches = max(1, n_samples // batch)
    assert len(zlog) == n_batches * (n_steps + 1) and len(slog) == n_batches * n_steps
    out = {"timesteps": model.noise_scheduler.timesteps.numpy().copy(),
           "step_size": np.float32(model.noise_scheduler.step_size.item()),
           "final": X.numpy(), "n_steps": np.int64(n_steps), "batch": np.int64(batch),
           "n_batches": np.int64(n_batches)}
    priors, noises = [], []
    for b in range(n_batches):
        chunk = zlog[b * (n_steps + 1):(b + 1) * (n_steps + 1)]
        priors.append(chunk[0].numpy())
        noises.append(np.stack([z.numpy() for z in chunk[1:]]))
    out["prior_z"] = np.stack(priors)  # [n_batches, B, L, C]
    out["step_z"] = np.stack(noises)  # [n_batches, n_steps, B, L, C]
    keep = sorted(set(s for s in keep_steps if s < n_batches * n_steps))
    out["keep_steps"] = np.array(keep, dtype=np.int64)
    out["keep_X"] = np.stack([slog[s]["X"].numpy() for s in keep])
    out["keep_Xn"] = np.stack([slog[s]["Xn"].numpy() for s in keep])
    out["keep_score"] = np.stack([sclog[s].numpy() for s in keep])
    out["keep_t"] = np.array([slog[s]["t"] for s in keep], dtype=np.float64)
    out["nS"] = np.array([-1 if s["S"] is None else len(s["S"]) for s in slog], dtype=np.int64)
    if use_cache:
        c = model.cache
        # cached layers alias the FIRST cache ever created for this model;
        # each golden run uses a fresh model so model.cache is that cache.
        out["k_cache"] = c.k_cache_tensor.numpy().copy()
        out["v_cache"] = c.v_cache_tensor.numpy().copy()
        out["crf"] = c.crf_cache.numpy().copy()
        st = c.get_cache_stats()
        out["stats_keys"] = np.array(sorted(st))
        out["stats_vals"] = np.array([float(st[k]) for k in sorted(st)])
    return out


def model_level(ref, model, X, t, sets):
    """Teacher-forced ScoreModule.forward with explicit recompute sets."""
    model.enable_caching(K=5, R=10)
    model.eval()
    model.cache.reset()
    scores, crfs = [], []
    ts = torch.full((X.shape[0],), t, dtype=torch.float32)
    with torch.no_grad():
        for i, S in enumerate(sets):
            model.cache.current_step = i
            batch = ref.DiffusableBatch(X=X + 0.01 * i, y=None, timesteps=ts)
            s, crf = model(batch, recompute_tokens=set(S), step=i, return_crf=True)
            scores.append(s.numpy().copy())
            crfs.append(crf.numpy().copy())
    return np.stack(scores), np.stack(crfs), model.cache.k_cache_tensor.numpy().copy(), \
        model.cache.v_cache_tensor.numpy().copy()


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    ref = rh.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- boundary transforms ------------------------------------------------
    g = torch.Generator().manual_seed(7)
    d = {}
    for L, C in ((24, 40), (100, 3), (101, 3), (187, 1), (251, 4), (252, 5), (365, 13)):
        x = torch.randn(2, L, C, generator=g)
        d[f"x_{L}"] = x.numpy()
        d[f"dft_{L}"] = ref.dft(x).numpy()
        d[f"idft_{L}"] = ref.idft(x).numpy()
    np.savez_compressed(os.path.join(OUT, "fourier.npz"), **d)

    # ---- FreSca score filter (fresca.py) ---------------------------------------
    sys.path.insert(0, os.path.join(rh.REFERENCE_ROOT, "src"))
    from fdiff.utils.fresca import apply_fresca_to_score  # type: ignore

    g3 = torch.Generator().manual_seed(3)
    d = {}
    for L, C in ((187, 1), (100, 3), (24, 5)):
        x = torch.randn(6, L, C, generator=g3) * (1.0 + torch.arange(L).float()[None, :, None] / L)
        d[f"x_{L}"] = x.numpy()
        for strat in ("spatial", "energy"):
            for tag, kw in (("a", dict(low_scale=1.0, high_scale=1.5, cutoff_ratio=0.5, timestep=0.37, num_steps=1000)),
                            ("b", dict(low_scale=0.8, high_scale=1.2, cutoff_ratio=0.3)),
                            ("c", dict(low_scale=1.3, high_scale=0.9, cutoff_ratio=0.7, timestep=0.9, num_steps=20))):
                d[f"y_{L}_{strat}_{tag}"] = apply_fresca_to_score(x, cutoff_strategy=strat, **kw).numpy()
    np.savez_compressed(os.path.join(OUT, "fresca.npz"), **d)
    m = rh.build_reference_model(ref, seed=11, sde="vp", fourier_noise_scaling=True, max_len=16, n_channels=2,
                                 d_model=12, num_layers=2, n_head=2)
    zlog, slog = [], []
    smp = ref.DiffusionSampler(m, sample_batch_size=3, use_cache=True, cache_kwargs=dict(K=2, R=100), use_fresca=True,
                               fresca_low_scale=1.0, fresca_high_scale=1.5, fresca_cutoff_ratio=0.5,
                               fresca_cutoff_strategy="energy")
    with rh.record_randn(zlog), rh.record_steps(smp, slog):
        Xf = smp.sample(3, 12)
    np.savez_compressed(os.path.join(OUT, "tiny_fresca_cache.npz"), final=Xf.numpy(), prior_z=zlog[0].numpy(),
                        step_z=np.stack([z.numpy() for z in zlog[1:]]),
                        timesteps=m.noise_scheduler.timesteps.numpy(),
                        keep_X=np.stack([e["X"].numpy() for e in slog]),
                        keep_Xn=np.stack([e["Xn"].numpy() for e in slog]))

    # ---- tiny model: weights stored ------------------------------------------
    tiny = dict(max_len=16, n_channels=2, d_model=12, num_layers=2, n_head=2)
    for tag, sde, fns, skw in (("vp_freq", "vp", True, {}), ("ve_freq", "ve", True, {"sigma_max": 2.0}),
                               ("vp_time", "vp", False, {})):
        for cache in (False, True):
            m = rh.build_reference_model(ref, seed=11, sde=sde, fourier_noise_scaling=fns,
                                         sde_kwargs=skw, **tiny)
            sd = rh.state_dict_numpy(m)
            tr = run_trace(ref, m, batch=3, n_samples=6 if cache else 3, n_steps=105 if cache else 8,
                           use_cache=cache, cache_kwargs=dict(K=2, R=100),
                           keep_steps=[0, 1, 2, 7, 99, 100, 101, 104, 105, 106, 200, 209])
            name = f"tiny_{tag}_{'cache' if cache else 'nocache'}.npz"
            np.savez_compressed(os.path.join(OUT, name), **tr,
                                **({f"w::{k}": v for k, v in sd.items()} if (tag == "vp_freq" and not cache) else {}),
                                pos_after=m.pos_encoder.embedding.weight.detach().numpy().copy())
            print(name, "ok")

    # teacher-forced model-level calls incl. arbitrary (non-prefix) sets and the >80% branch
    m = rh.build_reference_model(ref, seed=11, sde="vp", fourier_noise_scaling=True, **tiny)
    X = torch.randn(3, 16, 2, generator=g)
    sets = [range(16), [], range(4), [], [1, 5, 6, 11], [], range(14), [], range(16), []]
    sc, crf, kc, vc = model_level(ref, m, X, 0.37, sets)
    np.savez_compressed(os.path.join(OUT, "tiny_model_level.npz"), X=X.numpy(), t=np.float64(0.37),
                        sets=np.array([",".join(map(str, s)) for s in sets]), scores=sc, crf=crf,
                        k_cache=kc, v_cache=vc)
    # pure-cache call before any store (cache tensor never allocated -> zero K/V)
    m = rh.build_reference_model(ref, seed=11, sde="vp", fourier_noise_scaling=True, **tiny)
    sc, crf, _, _ = (lambda r: (r[0], r[1], None, None))(_cold(ref, m, X))
    np.savez_compressed(os.path.join(OUT, "tiny_cold_cache.npz"), X=X.numpy(), t=np.float64(0.37), score=sc, crf=crf)

    # ---- ECG default model: weights by seed + checksum -----------------------
    ecg = dict(max_len=187, n_channels=1, d_model=72, num_layers=10, n_head=12)
    m = rh.build_reference_model(ref, seed=42, sde="vp", fourier_noise_scaling=True, **ecg)
    keys, cs = checksums(rh.state_dict_numpy(m))
    tr = run_trace(ref, m, batch=2, n_samples=2, n_steps=103, use_cache=True, cache_kwargs=dict(K=5, R=100),
                   keep_steps=[0, 1, 2, 50, 99, 100, 101, 102])
    tr["k_cache"] = tr["k_cache"][[0, 9]]
    tr["v_cache"] = tr["v_cache"][[0, 9]]
    tr["crf"] = tr["crf"][[0, 9]]
    np.savez_compressed(os.path.join(OUT, "ecg_cache.npz"), **tr, ck_keys=keys, ck_vals=cs)
    print("ecg_cache ok")
    m = rh.build_reference_model(ref, seed=42, sde="vp", fourier_noise_scaling=True, **ecg)
    tr = run_trace(ref, m, batch=2, n_samples=2, n_steps=6, use_cache=False, cache_kwargs=None,
                   keep_steps=range(6))
    np.savez_compressed(os.path.join(OUT, "ecg_nocache.npz"), **tr, ck_keys=keys, ck_vals=cs)
    print("ecg_nocache ok")


def _cold(ref, model, X):
    model.enable_caching(K=5, R=10)
    model.eval()
    model.cache.reset()
    ts = torch.full((X.shape[0],), 0.37, dtype=torch.float32)
    with torch.no_grad():
        s, crf = model(ref.DiffusableBatch(X=X, y=None, timesteps=ts), recompute_tokens=set(), step=3,
                       return_crf=True)
    return s.numpy().copy(), crf.numpy().copy()


if __name__ == "__main__":
    main()
