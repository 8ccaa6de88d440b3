"""GPU parity tests: the CUDA path (through the C ABI) against the numpy oracle and the
golden vectors produced by the unmodified reference.  Tolerances: fp32 arithmetic
1e-5-class (stated per check), bf16 tensor-core arithmetic 1e-3-class on per-step states."""

import os

import numpy as np
import pytest
import torch

from ffd_testutil import (ECG, TINY, build_seeded_model, build_tiny_model, load_golden, rel, state_dict_checksums,
                          tiny_state_dict)
from oracle import fdiff_oracle as O

pytestmark = pytest.mark.gpu


def _precisions():
    out = ["fp32"]
    try:
        from fastfourierdiffusion_b200 import _lib

        if torch.cuda.is_available() and _lib.load().ffd_has_tensor_path():
            out.append("fp16")
    except Exception:
        pass
    return out


PRECS = ["fp32", "fp16"]
# per-precision tolerances: (score rel, one-step state rel, multi-step state rel)
# fp16 = tensor-core path (fp16 operands, fp32 accumulate): measured on B200 -- scores 1e-3 (well conditioned) to 3.6e-3
# (|X| ~ 126), one-step states 4e-5
TOL = {"fp32": dict(score=2e-5, step=2e-6, traj=5e-4), "fp16": dict(score=5e-3, step=2e-4, traj=2e-2)}


def _skip_unless(prec, model=None):
    if prec not in _precisions():
        pytest.skip("tensor path not available")
    if prec == "fp16" and model is not None and not model._tc_ok():
        pytest.skip("model shape not supported by the tcgen05 kernels")


def test_native_library_is_loaded():
    from fastfourierdiffusion_b200 import _lib

    lib = _lib.load()
    assert lib.ffd_version() >= 1
    maps = open("/proc/self/maps").read()
    assert "libffd_b200.so" in maps


# ---- boundary transforms -------------------------------------------------------------------
@pytest.mark.parametrize("L", [24, 100, 101, 187, 251, 252, 365])
def test_dft_idft_against_reference_vectors(L):
    import fastfourierdiffusion_b200 as F

    g = load_golden("fourier.npz")
    x = torch.from_numpy(g[f"x_{L}"])
    assert rel(F.dft(x).numpy(), g[f"dft_{L}"]) < 2e-6
    assert rel(F.idft(x).numpy(), g[f"idft_{L}"]) < 2e-6
    # reference tests/test_utils.py:36-51 (atol 1e-5)
    xc = x.cuda()
    assert torch.allclose(F.idft(F.dft(xc)), xc, atol=1e-5)
    assert torch.allclose(F.dft(F.idft(xc)), xc, atol=1e-5)


def test_dft_full_size_round_trip_and_parseval():
    import fastfourierdiffusion_b200 as F

    x = torch.randn(16384, 187, 1, device="cuda")
    y = F.dft(x)
    assert torch.allclose(F.idft(y), x, atol=2e-5)
    # orthonormal up to the packing: energy = re0^2 + 2*sum(others)
    w = torch.full((187,), 2.0, device="cuda")
    w[0] = 1.0
    e = (y[:, :, 0] ** 2 * w).sum(1)
    assert torch.allclose(e, (x[:, :, 0] ** 2).sum(1), rtol=1e-4)


def test_idft_fused_destandardise():
    import fastfourierdiffusion_b200 as F

    x = torch.randn(5, 24, 3, device="cuda")
    std, mean = torch.rand(24, 3, device="cuda") + 0.5, torch.randn(24, 3, device="cuda")
    assert torch.allclose(F.idft(x, std, mean), F.idft(x * std + mean), atol=1e-6)


# ---- small ops ------------------------------------------------------------------------------------
def test_sde_step_bit_exact_vs_oracle():
    import fastfourierdiffusion_b200 as F

    rng = np.random.default_rng(0)
    for kind, kw in (("vp", {}), ("ve", {"sigma_max": 2.0})):
        for L, C in ((187, 1), (24, 5)):
            s = (F.VPScheduler if kind == "vp" else F.VEScheduler)(fourier_noise_scaling=True, **kw)
            s.set_noise_scaling(L)
            s.set_timesteps(50)
            x, sc, z = (rng.standard_normal((7, L, C)).astype(np.float32) for _ in range(3))
            t = float(s.timesteps[13])
            got = s.step(torch.from_numpy(sc).cuda(), t, torch.from_numpy(x).cuda(),
                         noise=torch.from_numpy(z).cuda()).prev_sample.cpu().numpy()
            cfg = O.SDEConfig(kind=kind, fourier_noise_scaling=True, **kw)
            want = O.sde_step(cfg, O.noise_scaling(L, True), sc, t, x, z, np.float32(s.step_size.item()))
            assert np.array_equal(got, want)


def test_prior_bit_exact_and_philox_stream():
    import fastfourierdiffusion_b200 as F
    from fastfourierdiffusion_b200 import _lib

    s = F.VEScheduler(sigma_max=2.0, fourier_noise_scaling=True)
    s.set_noise_scaling(24)
    z = np.random.default_rng(1).standard_normal((3, 24, 2)).astype(np.float32)
    got = s.prior_from_noise(torch.from_numpy(z)).cpu().numpy()
    want = O.prior_sample(O.SDEConfig("ve", sigma_max=2.0, fourier_noise_scaling=True), O.noise_scaling(24, True), z)
    assert np.array_equal(got, want)
    out = torch.empty(10001, device="cuda")
    _lib.check(_lib.load().ffd_philox_normal(1234567890123, 77, _lib.ptr(out), out.numel(), _lib.stream_handle()))
    want = O.philox_normal(1234567890123, 77, 10001)
    assert np.abs(out.cpu().numpy() - want).max() < 1e-3  # fp32 Box-Muller vs the fp64 restatement
    big = torch.empty(1 << 22, device="cuda")
    _lib.check(_lib.load().ffd_philox_normal(5, 1, _lib.ptr(big), big.numel(), _lib.stream_handle()))
    assert abs(big.mean().item()) < 3e-3 and abs(big.std().item() - 1) < 3e-3


def test_time_embedding_and_posemb_renorm():
    m = build_tiny_model()
    sd = tiny_state_dict()
    ts = torch.linspace(1.0, 1e-5, 9)
    got = m.time_embedding(ts).cpu().numpy()
    for i, t in enumerate(ts.tolist()):
        want = O.time_embedding(t, sd["time_encoder.W"], sd["time_encoder.dense.weight"],
                                sd["time_encoder.dense.bias"], 12)
        assert np.abs(got[i] - want).max() < 2e-5  # |2 pi t W| ~ 1e2 rad: sin/cos differ by ulps of the argument
    m.device_weights()
    g = load_golden("tiny_vp_freq_nocache.npz")
    assert np.abs(m.pos_encoder.embedding.weight.detach().cpu().numpy() - g["pos_after"]).max() < 1e-6


# ---- score model ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", PRECS)
def test_tiny_model_level_sets(prec):
    """Every branch of the cached layer incl. non-prefix sets, against the reference's outputs."""
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model(precision=prec)
    _skip_unless(prec, m)
    g = load_golden("tiny_model_level.npz")
    m.enable_caching(K=5, R=10)
    m.cache.reset()
    X = torch.from_numpy(g["X"]).cuda()
    ts = torch.full((3,), float(g["t"]), dtype=torch.float32, device="cuda")
    tol = TOL[prec]["score"]
    for i, s in enumerate(g["sets"]):
        S = {int(v) for v in str(s).split(",") if v != ""}
        m.cache.current_step = i
        score, crf = m(F.DiffusableBatch(X=X + 0.01 * i, timesteps=ts), recompute_tokens=S, step=i, return_crf=True)
        assert rel(score.cpu().numpy(), g["scores"][i]) < tol, i
        assert rel(crf.cpu().numpy(), g["crf"][i]) < tol, i
    assert rel(m.cache.k_cache_tensor.cpu().numpy(), g["k_cache"]) < tol
    assert rel(m.cache.v_cache_tensor.cpu().numpy(), g["v_cache"]) < tol


@pytest.mark.parametrize("prec", PRECS)
def test_tiny_cold_cache_and_stock(prec):
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model(precision=prec)
    _skip_unless(prec, m)
    g = load_golden("tiny_cold_cache.npz")
    X = torch.from_numpy(g["X"]).cuda()
    ts = torch.full((3,), float(g["t"]), dtype=torch.float32, device="cuda")
    # stock path (no cache) vs the oracle
    w = O.ModelWeights.from_state_dict(tiny_state_dict(), 2)
    want = O.OracleScoreModel(w, 16, 2).forward(g["X"], float(g["t"]))
    got = m(F.DiffusableBatch(X=X, timesteps=ts))
    assert rel(got.cpu().numpy(), want) < TOL[prec]["score"]
    m.enable_caching()
    m.cache.reset()
    score, crf = m(F.DiffusableBatch(X=X, timesteps=ts), recompute_tokens=set(), step=3, return_crf=True)
    assert rel(score.cpu().numpy(), g["score"]) < TOL[prec]["score"]
    assert rel(crf.cpu().numpy(), g["crf"]) < TOL[prec]["score"]
    assert m.cache.k_cache_tensor is None  # nothing stored yet, as in the reference


def _run_trace(model, g, use_cache, cache_kwargs, noise_steps=None):
    import fastfourierdiffusion_b200 as F

    n = int(g["n_steps"])
    B = int(g["batch"])
    sampler = F.DiffusionSampler(model, B, use_cache=use_cache, cache_kwargs=cache_kwargs)
    model.noise_scheduler.set_timesteps(n)
    assert np.array_equal(model.noise_scheduler.timesteps.numpy(), g["timesteps"])
    outs, gs = [], 0
    for b in range(int(g["n_batches"])):
        X = model.noise_scheduler.prior_from_noise(torch.from_numpy(g["prior_z"][b]))
        if use_cache and b == 0:
            model.cache.reset()
            gs = 0
        X = sampler.run_steps(X, gs, batch_idx=b, noise=torch.from_numpy(g["step_z"][b]))
        gs += n
        outs.append(X.cpu().numpy())
    return sampler, np.concatenate(outs)


CASES = [("vp_freq", "vp", True, {}), ("ve_freq", "ve", True, {"sigma_max": 2.0}), ("vp_time", "vp", False, {})]


@pytest.mark.parametrize("tag,kind,fns,skw", CASES)
@pytest.mark.parametrize("cache", [False, True])
@pytest.mark.parametrize("prec", PRECS)
def test_tiny_sampler_trace(tag, kind, fns, skw, cache, prec):
    """Whole sampling runs (two sampler batches with the cache on) with the reference's noise."""
    m = build_tiny_model(kind, fns, precision=prec, **skw)
    _skip_unless(prec, m)
    g = load_golden(f"tiny_{tag}_{'cache' if cache else 'nocache'}.npz")
    sampler, out = _run_trace(m, g, cache, dict(K=2, R=100))
    assert rel(out, g["final"]) < TOL[prec]["traj"]
    if cache:
        c = m.cache
        assert rel(c.k_cache_tensor.cpu().numpy(), g["k_cache"]) < TOL[prec]["traj"] * 4
        assert rel(c.v_cache_tensor.cpu().numpy(), g["v_cache"]) < TOL[prec]["traj"] * 4
        assert rel(c.crf_cache.cpu().numpy(), g["crf"]) < TOL[prec]["traj"] * 4
        st = c.get_cache_stats()
        ref = dict(zip(g["stats_keys"], g["stats_vals"]))
        for k in ref:  # decisions and counters match exactly
            assert st[k] == pytest.approx(ref[k], rel=1e-12), k


@pytest.mark.parametrize("prec", PRECS)
def test_tiny_teacher_forced_steps(prec):
    """Per-step parity (well conditioned): feed the reference's X_t and z_t, compare X_{t-1}."""
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model(precision=prec)
    _skip_unless(prec, m)
    g = load_golden("tiny_vp_freq_cache.npz")
    n = int(g["n_steps"])
    sampler = F.DiffusionSampler(m, 3, use_cache=True, cache_kwargs=dict(K=2, R=100))
    m.noise_scheduler.set_timesteps(n)
    m.cache.reset()
    keep = {int(k): i for i, k in enumerate(g["keep_steps"]) if k < n}
    z = torch.from_numpy(g["step_z"][0])
    X = m.noise_scheduler.prior_from_noise(torch.from_numpy(g["prior_z"][0]))
    for i in range(n):
        if i in keep:
            X = torch.from_numpy(g["keep_X"][keep[i]]).cuda()
        X = sampler.run_steps(X, 0, noise=z, step_begin=i, step_end=i + 1)
        if i in keep:
            assert rel(X.cpu().numpy(), g["keep_Xn"][keep[i]]) < TOL[prec]["step"], i


@pytest.mark.parametrize("prec", PRECS)
def test_ecg_default_model_cached_trace(prec):
    """Hydra-default model (d=72, 10 layers, 12 heads, L=187): full / pure-cache / mixed steps."""
    import fastfourierdiffusion_b200 as F

    g = load_golden("ecg_cache.npz")
    m = build_seeded_model(ECG, 42, precision=prec)
    _skip_unless(prec, m)
    keys, cs = state_dict_checksums(m)
    np.testing.assert_allclose(cs, g["ck_vals"], rtol=1e-12)
    m = m.to("cuda").eval()
    n = int(g["n_steps"])
    sampler = F.DiffusionSampler(m, 2, use_cache=True, cache_kwargs=dict(K=5, R=100))
    m.noise_scheduler.set_timesteps(n)
    m.cache.reset()
    z = torch.from_numpy(g["step_z"][0])
    # teacher-forced at the kept steps (0 full, 1,2,50,99 pure, 100 mixed, 101,102 pure)
    keep = {int(k): i for i, k in enumerate(g["keep_steps"])}
    X = m.noise_scheduler.prior_from_noise(torch.from_numpy(g["prior_z"][0]))
    assert np.array_equal(X.cpu().numpy(), g["keep_X"][0])
    for i in range(n):
        if i in keep:
            X = torch.from_numpy(g["keep_X"][keep[i]]).cuda()
        X = sampler.run_steps(X, 0, noise=z, step_begin=i, step_end=i + 1)
        if i in keep:
            assert rel(X.cpu().numpy(), g["keep_Xn"][keep[i]]) < TOL[prec]["step"], i
    assert rel(X.cpu().numpy(), g["final"]) < TOL[prec]["step"]
    c = m.cache
    tol = 1e-4 if prec == "fp32" else 1e-2
    assert rel(c.k_cache_tensor[[0, 9]].cpu().numpy(), g["k_cache"]) < tol
    assert rel(c.v_cache_tensor[[0, 9]].cpu().numpy(), g["v_cache"]) < tol
    crf = c.crf_cache[[0, 9]].cpu().numpy()
    if prec == "fp32":
        assert rel(crf, g["crf"]) < tol
    else:
        # The CRF kept at the end is the one of the mixed step 100, where the random-init reverse SDE has driven
        # |X| to ~470: un-normalised layer-0 scores are in the hundreds, the softmax is nearly
        # one-hot, and bf16 operands flip the winning key for a few tokens (O(1) row errors at layer 0, damped by
        # the later LayerNorms).  So: the bulk of the rows at bf16 accuracy, the outliers bounded.
        err = np.abs(crf - g["crf"]).max(axis=-1) / np.abs(g["crf"]).max()
        print(f"ecg-vp [fp16] CRF of the ill-conditioned mixed step (|X| ~ 470): row error median {np.median(err):.2e}, "
              f"90th percentile {np.percentile(err, 90):.2e}, max {err.max():.2e}  (the well-conditioned counterpart, "
              "with no allowance, is test_ecg_ve_well_conditioned_trace)")
        assert np.percentile(err, 90) < 2e-3 and err.max() < 5e-2, (np.percentile(err, 90), err.max())
    st = c.get_cache_stats()
    ref = dict(zip(g["stats_keys"], g["stats_vals"]))
    for k in ref:
        assert st[k] == pytest.approx(ref[k], rel=1e-12), k


# tensor-core arithmetic (fp16 operands) on a WELL-CONDITIONED trace: score and CRF within 2.5e-3 (max-norm relative),
# K/V rows within 2e-3, per-step states within 2e-4 (the north star asks 1e-3) -- every row, no percentile allowance
TIGHT = {"fp32": dict(score=2e-5, crf=2e-5, kv=1e-4, step=2e-6), "fp16": dict(score=2.5e-3, crf=2.5e-3, kv=2e-3, step=2e-4)}


@pytest.mark.parametrize("prec", PRECS)
def test_ecg_ve_well_conditioned_trace(prec):
    """Hydra-default ECG model under the reference's VE-SDE (sigma_max = 2: |X| stays O(1) over the whole run), E2-CRF
    K = 5, R = 100: full step 0, pure-cache steps 1 / 2 / 50 / 99 / 101 / 103 and the MIXED step 100, teacher-forced
    against the unmodified reference (oracle/make_golden_ecg_ve.py).  The cache only changes at steps 0 and 100, so
    running the kept steps in order reproduces its state exactly.  Measured errors are printed."""
    import fastfourierdiffusion_b200 as F

    g = load_golden("ecg_ve_cache.npz")
    m = build_seeded_model(ECG, 42, kind="ve", precision=prec, sigma_min=0.01, sigma_max=2.0)
    _skip_unless(prec, m)
    keys, cs = state_dict_checksums(m)
    np.testing.assert_allclose(cs, g["ck_vals"], rtol=1e-12)
    m = m.to("cuda").eval()
    n = int(g["n_steps"])
    sampler = F.DiffusionSampler(m, 2, use_cache=True, cache_kwargs=dict(K=5, R=100))
    m.noise_scheduler.set_timesteps(n)
    assert np.array_equal(m.noise_scheduler.timesteps.numpy(), g["timesteps"])
    m.cache.reset()
    tol = TIGHT[prec]
    worst = dict(score=0.0, crf=0.0, step=0.0)
    z = g["step_z"][0]
    for j, step in enumerate(int(k) for k in g["keep_steps"]):
        X = torch.from_numpy(g["keep_X"][j]).cuda()
        t = float(g["keep_t"][j])
        ts = torch.full((2,), t, dtype=torch.float32, device="cuda")
        m.cache.current_step = step
        S = m.cache.determine_recompute_set(None, 0.1, step)
        assert len(S) == int(g["nS"][step])
        score, crf = m(F.DiffusableBatch(X=X, timesteps=ts), recompute_tokens=S, step=step, return_crf=True, t=t)
        e_s = rel(score.cpu().numpy(), g["keep_score"][j])
        Xn = m.noise_scheduler.step(model_output=score, timestep=t, sample=X,
                                    noise=torch.from_numpy(z[step]).cuda()).prev_sample
        e_x = rel(Xn.cpu().numpy(), g["keep_Xn"][j])
        worst["score"], worst["step"] = max(worst["score"], e_s), max(worst["step"], e_x)
        print(f"ecg-ve [{prec}] step {step:3d} |S|={len(S):3d}: score rel {e_s:.2e}, state rel {e_x:.2e}")
        assert e_s < tol["score"], (step, e_s)
        assert e_x < tol["step"], (step, e_x)
        m.cache.update_crf(crf, timestep=t)  # kept iff step % R == 0 (caching.py:474): steps 0 and 100
    c = m.cache
    e_k = rel(c.k_cache_tensor[[0, 9]].cpu().numpy(), g["k_cache"])
    e_v = rel(c.v_cache_tensor[[0, 9]].cpu().numpy(), g["v_cache"])
    e_c = rel(c.crf_cache[[0, 9]].cpu().numpy(), g["crf"])
    print(f"ecg-ve [{prec}] after the mixed step: K rel {e_k:.2e}, V rel {e_v:.2e}, CRF rel {e_c:.2e} (every row)")
    assert e_k < tol["kv"] and e_v < tol["kv"] and e_c < tol["crf"]


@pytest.mark.parametrize("prec", PRECS)
def test_ecg_mixed_branch_well_conditioned(prec):
    """The mixed branch (cached_transformer.py:259-305) of the default model at prior-sized inputs (|X| ~ 1, where
    bf16 arithmetic is well conditioned): prefix and non-prefix recompute sets against the oracle, score / CRF /
    stored K/V rows.  In bf16 mode this is the tensor-core path: per-sample K/V images with the rows outside the
    set patched in from the shared sample-0 images."""
    import fastfourierdiffusion_b200 as F

    m = build_seeded_model(ECG, 42, precision=prec)
    _skip_unless(prec, m)
    sd = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    m = m.to("cuda").eval()
    L = ECG["max_len"]
    rng = np.random.default_rng(11)
    B = 3
    m.enable_caching(K=5, R=10)
    m.cache.reset()
    om = O.OracleScoreModel(O.ModelWeights.from_state_dict(sd, ECG["n_head"]), L, ECG["n_channels"])
    store = O.KVStore(ECG["num_layers"], L)
    tol = TOL[prec]["score"]
    sets = [set(range(L)), set(range(10)), {3, 50, 186}, set(), set(range(0, L, 2))]
    for i, S in enumerate(sets):
        X = rng.normal(size=(B, L, 1)).astype(np.float32)
        t = 0.9 - 0.1 * i
        ts = torch.full((B,), t, dtype=torch.float32, device="cuda")
        m.cache.current_step = i
        score, crf = m(F.DiffusableBatch(X=torch.from_numpy(X).cuda(), timesteps=ts), recompute_tokens=S, step=i,
                       return_crf=True)
        want_s, want_c = om.forward_cached(X, t, sorted(S), store)
        assert rel(score.cpu().numpy(), want_s) < tol, (i, "score")
        assert rel(crf.cpu().numpy(), want_c) < tol, (i, "crf")
        own = m._kv_owner
        assert rel(own._k_store.cpu().numpy(), store.k) < tol, (i, "k")
        assert rel(own._v_store.cpu().numpy(), store.v) < tol, (i, "v")


@pytest.mark.parametrize("prec", PRECS)
def test_ecg_default_model_scores_and_free_run(prec):
    import fastfourierdiffusion_b200 as F

    g = load_golden("ecg_nocache.npz")
    m = build_seeded_model(ECG, 42, precision=prec)
    _skip_unless(prec, m)
    m = m.to("cuda").eval()
    m.noise_scheduler.set_timesteps(int(g["n_steps"]))
    for i, st in enumerate(g["keep_steps"]):
        X = torch.from_numpy(g["keep_X"][i]).cuda()
        ts = torch.full((2,), float(g["keep_t"][i]), dtype=torch.float32, device="cuda")
        score = m(F.DiffusableBatch(X=X, timesteps=ts))
        e_s = rel(score.cpu().numpy(), g["keep_score"][i])
        print(f"ecg-vp no cache [{prec}] step {int(st)}: score rel {e_s:.2e} (max |X| {float(X.abs().max()):.1f})")
        assert e_s < TOL[prec]["score"], st
    sampler, out = _run_trace(m, g, False, None)
    assert rel(out, g["final"]) < TOL[prec]["traj"]


# ---- the public sampler API ---------------------------------------------------------------------------------
def test_sampler_public_api_shapes_and_rng_parity():
    """reference tests/test_sampling.py:21-40 (shape contract) + same torch RNG consumption as
    the reference: CPU generator for the prior, device generator per step."""
    import fastfourierdiffusion_b200 as F

    for kind in ("vp", "ve"):
        torch.manual_seed(42)
        sched = (F.VPScheduler if kind == "vp" else F.VEScheduler)(fourier_noise_scaling=True)
        m = F.ScoreModule(n_channels=3, max_len=50, noise_scheduler=sched, precision="fp32").cuda()
        sched.set_noise_scaling(50)
        sampler = F.DiffusionSampler(score_model=m, sample_batch_size=12)
        samples = sampler.sample(num_samples=48, num_diffusion_steps=10)
        assert samples.shape == (48, 50, 3) and samples.device.type == "cpu" and samples.dtype == torch.float32
        assert torch.isfinite(samples).all()
    # remainder dropped / single short batch (sampler.py:127,143-146)
    assert sampler.sample(num_samples=30, num_diffusion_steps=3).shape == (24, 50, 3)
    assert sampler.sample(num_samples=5, num_diffusion_steps=3).shape == (5, 50, 3)
    # torch-noise mode reproduces a manual loop that draws randn_like per step
    torch.manual_seed(7)
    a = sampler.sample(12, 6)
    torch.manual_seed(7)
    sched.set_timesteps(6)
    X = sched.prior_sampling((12, 50, 3))
    for t in sched.timesteps:
        ts = torch.full((12,), t.item(), dtype=torch.float32, device="cuda")
        X = sampler.reverse_diffusion_step(F.DiffusableBatch(X=X, timesteps=ts))
    assert torch.allclose(a, X.cpu(), rtol=1e-5, atol=1e-5)


def test_philox_sampler_is_deterministic_and_batch_invariant():
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model()
    s = F.DiffusionSampler(m, 64, use_cache=True, cache_kwargs=dict(K=2, R=100), noise="philox", seed=3)
    a = s.sample(64, 20)
    b = s.sample(64, 20)
    assert torch.equal(a, b)
    assert torch.isfinite(a).all() and a.std() > 0
    s2 = F.DiffusionSampler(m, 64, use_cache=True, cache_kwargs=dict(K=2, R=100), noise="philox", seed=4)
    assert not torch.equal(a, s2.sample(64, 20))


def test_error_conventions():
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model()
    ts = torch.full((3,), 0.5, device="cuda")
    with pytest.raises(AssertionError):  # score_models.py:87-90
        m(F.DiffusableBatch(X=torch.zeros(3, 15, 2, device="cuda"), timesteps=ts))
    with pytest.raises(AssertionError):  # score_models.py:93
        m(F.DiffusableBatch(X=torch.zeros(3, 16, 2, device="cuda"), timesteps=ts[:2]))
    s = F.DiffusionSampler(m, 3)
    with pytest.raises(AssertionError):  # sampler.py:59-60
        s.reverse_diffusion_step(F.DiffusableBatch(X=torch.zeros(3, 16, 2, device="cuda"),
                                                   timesteps=torch.tensor([0.1, 0.2, 0.3], device="cuda")))
    with pytest.raises(IndexError):  # one step has no step size (sde.py:64)
        s.sample(3, 1)
    with pytest.raises(TypeError):  # README's random_probe_ratio is not a cache kwarg (caching.py:28-47)
        F.DiffusionSampler(m, 3, use_cache=True, cache_kwargs={"random_probe_ratio": 0.1})


# ---- full-size properties (BASELINE config 2 shapes) ---------------------------------------------------------
@pytest.mark.parametrize("prec", PRECS)
def test_full_batch_properties(prec):
    """B=16384 x L=187: (1) no-cache outputs are per-sample functions (a permuted batch gives the
    permuted result), (2) with the cache on every sample sees sample 0's K/V: replicating sample 0
    through the batch reproduces sample 0's trajectory everywhere."""
    import fastfourierdiffusion_b200 as F

    m = build_seeded_model(ECG, 42, precision=prec)
    _skip_unless(prec, m)
    m = m.to("cuda").eval()
    B = 16384
    sched = m.noise_scheduler
    sched.set_timesteps(1000)
    torch.manual_seed(0)
    X = sched.prior_from_noise(torch.randn(B, 187, 1))
    ts = torch.full((B,), float(sched.timesteps[3]), device="cuda")
    s1 = m(F.DiffusableBatch(X=X, timesteps=ts))
    perm = torch.randperm(B, device="cuda")
    s2 = m(F.DiffusableBatch(X=X[perm].contiguous(), timesteps=ts))
    assert torch.isfinite(s1).all()
    assert torch.allclose(s1[perm], s2, rtol=0, atol=0 if prec == "fp32" else 1e-6 * float(s1.abs().max()) + 1e-6)
    small = m(F.DiffusableBatch(X=X[:4].contiguous(), timesteps=ts[:4]))
    assert rel(s1[:4].cpu().numpy(), small.cpu().numpy()) < 1e-6
    sampler = F.DiffusionSampler(m, B, use_cache=True, noise="philox")
    m.cache.reset()
    Xc = X[:1].expand(B, -1, -1).contiguous()
    # three steps (full, pure, pure) with identical per-sample noise
    nz = torch.randn(3, 1, 187, 1, device="cuda").expand(3, B, 187, 1).contiguous()
    Xc = sampler.run_steps(Xc, 0, noise=nz, step_begin=0, step_end=3)
    assert torch.isfinite(Xc).all()
    assert rel(Xc.cpu().numpy(), Xc[:1].expand(B, -1, -1).cpu().numpy()) < (1e-6 if prec == "fp32" else 1e-5)


def test_entry_points_run(tmp_path):
    """cmd/sample.py and cmd/benchmark_cache.py with the reference's key names (tiny settings)."""
    import subprocess
    import sys

    from ffd_testutil import ROOT

    out = tmp_path / "o"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "cmd", "sample.py"), "num_samples=8",
                        "num_diffusion_steps=5", "use_cache=true", "cache_kwargs={'K': 5, 'R': 10}",
                        "sampler.sample_batch_size=4", "score_model.num_layers=2", f"output_dir={out}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    X = torch.load(out / "samples.pt")
    assert X.shape == (8, 187, 1) and torch.isfinite(X).all()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "cmd", "benchmark_cache.py"), "num_samples=2",
                        "num_diffusion_steps=4", "score_model.num_layers=1", f"output_dir={out}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (out / "benchmark_results.csv").exists()
    # cmd/ablation_cache.py on the time-domain model (BASELINE.json config 5: no DFT, G = 1)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "cmd", "ablation_cache.py"), "num_samples=4",
                        "num_diffusion_steps=6", "fourier_transform=false", "sampler.sample_batch_size=2",
                        "score_model.num_layers=1", f"output_dir={out}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    import json

    res = json.load(open(out / "ablation_results.json"))
    assert list(res) == ["Baseline", "E2-CRF (full)", "No event trigger", "No error feedback", "No energy weighting",
                         "Naive caching"]
    assert all(v["finite"] for v in res.values())
    # first cache of the model (K=5, R=10: refresh every 500 steps): step 0 recomputes, the 11 later global steps
    # (two sampler batches share the cache) are cache hits for all L tokens of the layer.  Later configurations
    # report zero counters, like the reference: the cached layers keep the FIRST cache object.
    st = res["E2-CRF (full)"]["cache_stats"]
    assert st["recompute_count"] == 187 and st["cache_hit_count"] == 187 * (2 * 6 - 1)
    assert res["No error feedback"]["cache_stats"]["recompute_count"] == 0


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("shape", [dict(max_len=200, n_channels=3, d_model=24, num_layers=2, n_head=4),
                                   dict(max_len=37, n_channels=5, d_model=60, num_layers=2, n_head=12),
                                   # BASELINE.json configs 3 and 4: NASA-charge, NASDAQ, USDroughts (default model)
                                   dict(max_len=251, n_channels=4, d_model=72, num_layers=2, n_head=12),
                                   dict(max_len=252, n_channels=5, d_model=72, num_layers=2, n_head=12),
                                   dict(max_len=365, n_channels=13, d_model=72, num_layers=2, n_head=12),
                                   # head_dim 12: one head per Q tile, row sum in slot 15, K/V images of the stock
                                   # branch from the fp32 projection + pack kernel; odd head count
                                   dict(max_len=50, n_channels=2, d_model=48, num_layers=2, n_head=4),
                                   dict(max_len=70, n_channels=1, d_model=36, num_layers=2, n_head=3),
                                   # three key chunks of 192: the padded keys 512..575 lie beyond the four row tiles
                                   dict(max_len=385, n_channels=2, d_model=24, num_layers=1, n_head=4)])
def test_other_shapes_against_oracle(prec, shape):
    """Other geometries of the tensor-core attention kernel -- several key chunks per head (L > 192: 2 x 128 or
    2 x 192 keys), three heads per Q tile (d=60, hd=5), one 64-key chunk, multi-channel -- every branch against the
    numpy oracle with seeded weights (no reference run needed)."""
    import fastfourierdiffusion_b200 as F

    m = build_seeded_model(shape, 5, precision=prec)
    if prec not in _precisions():
        pytest.skip("tensor path not available")
    if prec == "fp16" and not m._tc_ok():
        pytest.skip("model shape not supported by the tcgen05 kernels")
    sd = {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    m = m.to("cuda").eval()
    L, C = shape["max_len"], shape["n_channels"]
    w = O.ModelWeights.from_state_dict(sd, shape["n_head"])
    om = O.OracleScoreModel(w, L, C)
    rng = np.random.default_rng(1)
    X = rng.standard_normal((3, L, C)).astype(np.float32)
    t = 0.61
    ts = torch.full((3,), t, dtype=torch.float32, device="cuda")
    Xd = torch.from_numpy(X).cuda()
    tol = TOL[prec]["score"]
    got = m(F.DiffusableBatch(X=Xd, timesteps=ts))
    assert rel(got.cpu().numpy(), om.forward(X, t)) < tol
    m.enable_caching(K=3, R=10)
    m.cache.reset()
    store = O.KVStore(shape["num_layers"], L)
    for i, S in enumerate([range(L), [], range(6), [], [2, 9, 30], []]):
        score, crf = m(F.DiffusableBatch(X=Xd + 0.02 * i, timesteps=ts), recompute_tokens=set(S), step=i,
                       return_crf=True)
        want, wcrf = om.forward_cached(X + np.float32(0.02 * i), t, list(S), store)
        assert rel(score.cpu().numpy(), want) < tol, i
        assert rel(crf.cpu().numpy(), wcrf) < tol, i


@pytest.mark.parametrize("L", [187, 100, 24])
@pytest.mark.parametrize("strat", ["spatial", "energy"])
def test_fresca_filter_against_reference_vectors(L, strat):
    """apply_fresca_to_score (fresca.py:213-268): both cut-off strategies, odd/even L, time-dependent h."""
    import fastfourierdiffusion_b200 as F

    g = load_golden("fresca.npz")
    kws = {"a": dict(low_scale=1.0, high_scale=1.5, cutoff_ratio=0.5, timestep=0.37, num_steps=1000),
           "b": dict(low_scale=0.8, high_scale=1.2, cutoff_ratio=0.3),
           "c": dict(low_scale=1.3, high_scale=0.9, cutoff_ratio=0.7, timestep=0.9, num_steps=20)}
    x = torch.from_numpy(g[f"x_{L}"]).cuda()
    for tag, kw in kws.items():
        got = F.apply_fresca_to_score(x, cutoff_strategy=strat, **kw)
        assert rel(got.cpu().numpy(), g[f"y_{L}_{strat}_{tag}"]) < 5e-6, (L, strat, tag)
    assert F.apply_fresca_to_score(x, 1.0, 1.0) is x  # no-op short cut (fresca.py:134-135)


@pytest.mark.parametrize("prec", PRECS)
def test_tiny_sampler_trace_with_fresca(prec):
    """DiffusionSampler(use_fresca=True) with the cache on, teacher-forced per step against the reference."""
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model(precision=prec)
    _skip_unless(prec, m)
    g = load_golden("tiny_fresca_cache.npz")
    s = F.DiffusionSampler(m, 3, use_cache=True, cache_kwargs=dict(K=2, R=100), use_fresca=True,
                           fresca_low_scale=1.0, fresca_high_scale=1.5, fresca_cutoff_ratio=0.5,
                           fresca_cutoff_strategy="energy")
    n = g["step_z"].shape[0]
    s._num_diffusion_steps = n
    m.noise_scheduler.set_timesteps(n)
    m.cache.reset()
    for i, t in enumerate(m.noise_scheduler.timesteps.tolist()):
        X = torch.from_numpy(g["keep_X"][i]).cuda()
        ts = torch.full((3,), t, dtype=torch.float32, device="cuda")
        m.cache.current_step = i
        S = m.cache.determine_recompute_set(None, 0.1, i)
        Xn = s.reverse_diffusion_step(F.DiffusableBatch(X=X, timesteps=ts), step=i, recompute_tokens=S,
                                      noise=torch.from_numpy(g["step_z"][i]).cuda())
        assert rel(Xn.cpu().numpy(), g["keep_Xn"][i]) < TOL[prec]["step"], i
    # free run through the FUSED loop (ffd_sample_steps with ffd_fresca_args: what sample() calls) ...
    X0 = m.noise_scheduler.prior_from_noise(torch.from_numpy(g["prior_z"]))
    m.cache.reset()
    from fastfourierdiffusion_b200 import _lib

    lib = _lib.load()
    n0 = lib.ffd_launch_count()
    out = s.run_steps(X0, 0, noise=torch.from_numpy(g["step_z"]).cuda())
    assert lib.ffd_launch_count() > n0
    err = rel(out.cpu().numpy(), g["final"])
    print(f"fresca fused free run [{prec}]: rel {err:.2e}")
    assert err < TOL[prec]["traj"]
    # ... and through the step-by-step form of the public API: same kernels, same values
    m.cache.reset()
    out2 = s.run_steps_unfused(X0, 0, noise=torch.from_numpy(g["step_z"]).cuda())
    assert rel(out2.cpu().numpy(), out.cpu().numpy()) < (1e-6 if prec == "fp32" else TOL[prec]["traj"])


@pytest.mark.parametrize("tag", ["tiny", "ecg", "sat", "nasa"])
def test_event_intensity_against_reference_vectors(tag):
    """E2CRFCache.compute_event_intensity (caching.py:524-559) through ffd_event_intensity; the device variant
    returns the value without a host sync."""
    import fastfourierdiffusion_b200 as F

    g = load_golden("event_intensity.npz")
    prev, cur, tau = g[f"{tag}_prev"], g[f"{tag}_cur"], float(g[f"{tag}_tau"])
    NL, L, d = prev.shape
    c = F.E2CRFCache(num_layers=NL, max_len=L, device=torch.device("cuda"), tau_0=tau)
    c.attach_stores(1, d, d)
    assert c.compute_event_intensity(torch.from_numpy(cur).cuda(), 0) == 1.0
    assert c.compute_event_intensity(torch.from_numpy(cur).cuda(), 5) == 0.1
    c.current_step = 0
    c.update_crf(torch.from_numpy(prev).cuda())
    want = float(g[f"{tag}_val"])
    got = c.compute_event_intensity(torch.from_numpy(cur).cuda(), 3)
    assert abs(got - want) <= 2e-6 * max(1.0, abs(want))
    dev = c.event_intensity_device(torch.from_numpy(cur).cuda())
    assert dev.is_cuda and dev.shape == (2,)
    assert abs(float(dev[1]) - want) <= 2e-6 * max(1.0, abs(want))
    assert abs(float(dev[0]) - O.event_intensity(cur, prev, 1e30) * 1e30) <= 1e-5 * float(dev[0])


@pytest.mark.parametrize("n,m,shape", [(1000, 777, (187, 1)), (4096, 4096, (24, 3)), (300, 1, (16, 2)),
                                       (16384, 5000, (187, 1))])
def test_wasserstein_metrics_against_oracle(n, m, shape):
    """SlicedWasserstein / MarginalWasserstein (metrics.py:101-216) on the device against the numpy restatement:
    same directions (host RNG), distances within 1e-4 relative (fp32 projections vs float64)."""
    import fastfourierdiffusion_b200 as F
    from oracle import metrics_oracle as MO

    rng = np.random.default_rng(5)
    L, C = shape
    orig = rng.normal(size=(n, L, C)).astype(np.float32)
    other = (rng.normal(size=(m, L, C)) * 1.3 + 0.4 * np.sin(np.arange(L))[None, :, None]).astype(np.float32)
    ndir = 16 if n > 5000 else 40
    got = F.SlicedWasserstein(orig, random_seed=42, num_directions=ndir, save_all_distances=True)(torch.from_numpy(other))
    want, wd = MO.sliced_wasserstein(orig, other, 42, ndir)
    assert np.allclose(got["sliced_wasserstein_all"], wd, rtol=1e-4, atol=1e-5)
    assert abs(got["sliced_wasserstein_mean"] - want["sliced_wasserstein_mean"]) < 1e-4 * want["sliced_wasserstein_mean"]
    assert abs(got["sliced_wasserstein_max"] - want["sliced_wasserstein_max"]) < 1e-4 * want["sliced_wasserstein_max"]
    if n <= 5000:
        gm = F.MarginalWasserstein(orig, random_seed=42, save_all_distances=True)(other)
        wm, wdm = MO.marginal_wasserstein(orig, other, 42)
        assert np.allclose(gm["marginal_wasserstein_all"], wdm, rtol=1e-4, atol=1e-5)
        assert abs(gm["marginal_wasserstein_mean"] - wm["marginal_wasserstein_mean"]) < 1e-4 * wm["marginal_wasserstein_mean"]
        # standardise + single direction helpers
        w = F.WassersteinDistances(orig, other, normalisation="standardise", seed=7)
        wo = MO.WassersteinDistances(orig.reshape(n, -1), other.reshape(m, -1), normalisation="standardise", seed=7)
        assert np.allclose(w.sliced_distances(6), wo.sliced_distances(6), rtol=1e-4, atol=1e-5)
        assert abs(w.feature_distance(3) - wo.feature_distance(3)) < 1e-4
    if n == 1000:
        base = F.SlicedWasserstein(orig, random_seed=1, num_directions=8).baseline_metrics
        assert set(base) == {"sliced_wasserstein_mean_self", "sliced_wasserstein_max_self",
                             "sliced_wasserstein_mean_dummy", "sliced_wasserstein_max_dummy"}
        assert 0 < base["sliced_wasserstein_mean_self"] < base["sliced_wasserstein_mean_dummy"]
    with pytest.raises(Exception):
        F.WassersteinDistances(np.zeros((40000, 4), np.float32), np.zeros((10, 4), np.float32))


@pytest.mark.parametrize("L", [24, 187, 101, 252])
def test_spectral_density_against_reference_vectors(L):
    """spectral_density (fourier.py:97-131) through ffd_dft_apply + ffd_spectral_density."""
    import fastfourierdiffusion_b200 as F

    g = load_golden("spectral.npz")
    x = torch.from_numpy(g[f"x_{L}"]).cuda()
    got = F.spectral_density(x)
    assert got.shape == g[f"sd_{L}"].shape
    assert np.allclose(got.cpu().numpy(), g[f"sd_{L}"], rtol=2e-4, atol=2e-5)
    got2 = F.spectral_density(F.dft(x), apply_dft=False)
    assert np.allclose(got2.cpu().numpy(), g[f"sdf_{L}"], rtol=2e-4, atol=2e-5)
    assert F.spectral_density(x.cpu()).device.type == "cpu"  # returned on the caller's device, like dft


def test_metric_collection_against_oracle():
    """MetricCollection (metrics.py:28-98): time / freq / spectral keys, values against the numpy restatement."""
    from functools import partial

    import fastfourierdiffusion_b200 as F
    from oracle import metrics_oracle as MO

    rng = np.random.default_rng(9)
    orig = rng.normal(size=(600, 24, 2)).astype(np.float32)
    other = (rng.normal(size=(500, 24, 2)) * 1.2 + 0.2).astype(np.float32)
    mc = F.MetricCollection([partial(F.SlicedWasserstein, random_seed=42, num_directions=12),
                             partial(F.MarginalWasserstein, random_seed=42)], original_samples=orig,
                            include_baselines=True, include_spectral_density=True)
    res = mc(other)
    assert list(res) == sorted(res)
    for k in ("time_sliced_wasserstein_mean", "freq_sliced_wasserstein_max", "time_marginal_wasserstein_mean",
              "freq_marginal_wasserstein_mean_self", "time_sliced_wasserstein_mean_dummy",
              "spectral_marginal_wasserstein_mean", "spectral_marginal_wasserstein_all"):
        assert k in res, k
    o64, x64 = orig.astype(np.float64), other.astype(np.float64)
    want_t, _ = MO.sliced_wasserstein(o64, x64, 42, 12)
    want_f, _ = MO.sliced_wasserstein(O.dft(o64), O.dft(x64), 42, 12)
    want_s, _ = MO.marginal_wasserstein(O.spectral_density(o64), O.spectral_density(x64), 42)
    assert abs(res["time_sliced_wasserstein_mean"] - want_t["sliced_wasserstein_mean"]) < 1e-4
    assert abs(res["freq_sliced_wasserstein_mean"] - want_f["sliced_wasserstein_mean"]) < 1e-4
    assert abs(res["spectral_marginal_wasserstein_mean"] - want_s["marginal_wasserstein_mean"]) < 2e-4


def test_sample_time_domain_equals_idft_of_sample():
    """sample_time_domain = sample() + the de-standardise / idft tail of cmd/sample.py:107-113 on the device."""
    import fastfourierdiffusion_b200 as F

    m = build_tiny_model()
    L, Cc = TINY["max_len"], TINY["n_channels"]
    rng = np.random.default_rng(2)
    std = torch.from_numpy(rng.uniform(0.5, 2.0, size=(L, Cc)).astype(np.float32))
    mean = torch.from_numpy(rng.normal(size=(L, Cc)).astype(np.float32))
    outs = []
    for fn in ("sample", "sample_time_domain"):
        s = F.DiffusionSampler(m, 4, use_cache=True, cache_kwargs=dict(K=2, R=100), noise="philox", seed=9)
        if fn == "sample":
            outs.append(F.idft(s.sample(8, 12), std, mean))
        else:
            outs.append(s.sample_time_domain(8, 12, std, mean))
    assert outs[0].shape == outs[1].shape == (8, L, Cc) and outs[1].device.type == "cpu"
    assert torch.equal(outs[0], outs[1])
    want = O.idft(F.DiffusionSampler(m, 4, use_cache=True, cache_kwargs=dict(K=2, R=100), noise="philox",
                                     seed=9).sample(8, 12).numpy().astype(np.float64) * std.numpy() + mean.numpy())
    assert rel(outs[1].numpy(), want) < 2e-5


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("noise", ["philox", "torch"])
def test_graph_replay_equals_direct_launches(prec, noise, monkeypatch):
    """The captured-and-replayed step loop (ffd_sample_args.graph: two pure-cache steps per CUDA graph, per-step scalars
    read from device memory) runs the same kernels on the same data as step-by-step launches: identical bits, for
    in-kernel Philox noise and for injected (torch) noise, across several sampler batches that share one cache."""
    import fastfourierdiffusion_b200 as F
    from fastfourierdiffusion_b200 import _lib

    m = build_seeded_model(ECG, 42, precision=prec)
    _skip_unless(prec, m)
    m = m.to("cuda").eval()
    lib = _lib.load()
    outs, launches = {}, {}
    for mode in ("0", "1"):
        monkeypatch.setenv("FFD_GRAPH", mode)
        s = F.DiffusionSampler(m, 4, use_cache=True, cache_kwargs=dict(K=5, R=10), noise=noise, seed=9)
        torch.manual_seed(5)
        n0 = lib.ffd_launch_count()
        outs[mode] = s.sample(12, 45)  # 3 sampler batches x 45 steps: CRF-store steps every 10 break the runs
        launches[mode] = lib.ffd_launch_count() - n0
        assert torch.isfinite(outs[mode]).all()
    assert torch.equal(outs["0"], outs["1"])
    # replays are counted kernel by kernel; the graph adds one 1-thread cursor kernel per replayed step
    assert launches["0"] <= launches["1"] <= launches["0"] + 3 * 45, launches
