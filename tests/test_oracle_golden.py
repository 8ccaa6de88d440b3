"""Pin the numpy oracle to vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""

import numpy as np
import pytest

from oracle import fdiff_oracle as O
from ffd_testutil import load_golden

TINY = dict(L=16, C=2, H=2)


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("L", [24, 100, 101, 187, 251, 252, 365])
def test_fourier_matrix_form(L):
    g = load_golden("fourier.npz")
    x = g[f"x_{L}"].astype(np.float64)
    assert rel(O.dft(x), g[f"dft_{L}"]) < 2e-6
    assert rel(O.idft(x), g[f"idft_{L}"]) < 2e-6
    # reference tests/test_utils.py:36-51 -- mutual inverses
    assert np.abs(O.idft(O.dft(x)) - x).max() < 1e-10
    assert np.abs(O.dft(O.idft(x)) - x).max() < 1e-10


def _sampler(weights, L, C, H, kind, fns, dtype, use_cache, K=2, R=100, sigma_max=50.0):
    w = O.ModelWeights.from_state_dict(weights, H, dtype=np.float32)
    model = O.OracleScoreModel(w, L, C, dtype=dtype)
    sde = O.SDEConfig(kind=kind, fourier_noise_scaling=fns, sigma_max=sigma_max)
    return O.OracleSampler(model, sde, use_cache=use_cache, K=K, R=R)


CASES = [("vp_freq", "vp", True, 50.0), ("ve_freq", "ve", True, 2.0), ("vp_time", "vp", False, 50.0)]


@pytest.mark.parametrize("tag,kind,fns,smax", CASES)
@pytest.mark.parametrize("cache", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tiny_sampler_trace(tiny_weights, tag, kind, fns, smax, cache, dtype):
    g = load_golden(f"tiny_{tag}_{'cache' if cache else 'nocache'}.npz")
    s = _sampler(tiny_weights, TINY["L"], TINY["C"], TINY["H"], kind, fns, dtype, cache, sigma_max=smax)
    n = int(g["n_steps"])
    tr = O.OracleTrace()
    out = s.sample_batches(list(g["prior_z"]), [list(z) for z in g["step_z"]], n, trace=tr,
                           ts=g["timesteps"])
    # recompute decisions match exactly
    nS = [-1 if S is None else len(S) for S in tr.recompute]
    assert nS == list(g["nS"])
    for i, st in enumerate(g["keep_steps"]):
        assert rel(tr.states[st], g["keep_X"][i]) < 2e-4, st
        assert rel(tr.scores[st], g["keep_score"][i]) < 2e-3, st
    assert rel(out, g["final"]) < 5e-4
    if cache:
        assert rel(s.store.k, g["k_cache"]) < 1e-3
        assert rel(s.store.v, g["v_cache"]) < 1e-3
        assert rel(s.crf, g["crf"]) < 1e-3
        st = O.cache_stats(s.store, n - 1)
        ref = dict(zip(g["stats_keys"], g["stats_vals"]))
        for k in ("recompute_count", "cache_hit_count", "cache_hit_ratio", "cache_ratio"):
            assert st[k] == pytest.approx(ref[k], rel=1e-9), k


def test_tiny_teacher_forced_steps(tiny_weights):
    """Per-step (well-conditioned) parity: feed the reference's X_t, z_t."""
    g = load_golden("tiny_vp_freq_cache.npz")
    s = _sampler(tiny_weights, 16, 2, 2, "vp", True, np.float32, True)
    n = int(g["n_steps"])
    ts = g["timesteps"]
    dt = np.float32(ts[0] - ts[1])
    assert dt == g["step_size"]
    # replay batch 0 fully so the cache state is right, checking kept steps teacher-forced
    X = O.prior_sample(s.sde, s.G, g["prior_z"][0])
    keep = {int(k): i for i, k in enumerate(g["keep_steps"])}
    for i in range(n):
        S = O.determine_recompute_set(i, 2, 100, 16)
        if i in keep:
            Xin = g["keep_X"][keep[i]]
            assert float(g["keep_t"][keep[i]]) == pytest.approx(float(ts[i]), abs=0)
            Xn, sc = s.step(Xin, float(ts[i]), g["step_z"][0][i], dt, S, i)
            assert rel(sc, g["keep_score"][keep[i]]) < 5e-5
            assert rel(Xn, g["keep_Xn"][keep[i]]) < 2e-6
            X = Xn
        else:
            X, _ = s.step(X, float(ts[i]), g["step_z"][0][i], dt, S, i)


def test_tiny_model_level_sets(tiny_weights):
    g = load_golden("tiny_model_level.npz")
    w = O.ModelWeights.from_state_dict(tiny_weights, 2)
    model = O.OracleScoreModel(w, 16, 2, dtype=np.float32)
    store = O.KVStore(2, 16)
    for i, s in enumerate(g["sets"]):
        S = [int(v) for v in str(s).split(",") if v != ""]
        score, crf = model.forward_cached(g["X"] + np.float32(0.01 * i), float(g["t"]), S, store)
        assert rel(score, g["scores"][i]) < 5e-5, i
        assert rel(crf, g["crf"][i]) < 5e-5, i
    assert rel(store.k, g["k_cache"]) < 1e-5
    assert rel(store.v, g["v_cache"]) < 1e-5


def test_tiny_cold_cache_is_zero_kv(tiny_weights):
    g = load_golden("tiny_cold_cache.npz")
    w = O.ModelWeights.from_state_dict(tiny_weights, 2)
    model = O.OracleScoreModel(w, 16, 2, dtype=np.float32)
    score, crf = model.forward_cached(g["X"], float(g["t"]), [], O.KVStore(2, 16))
    assert rel(score, g["score"]) < 5e-5
    assert rel(crf, g["crf"]) < 5e-5


def test_pos_embedding_renorm(tiny_weights):
    g = load_golden("tiny_vp_freq_nocache.npz")
    got = O.posemb_renorm(tiny_weights["pos_encoder.embedding.weight"], np.sqrt(12.0), iters=8)
    assert np.abs(got - g["pos_after"]).max() < 1e-6


def test_linspace_close_to_torch():
    import torch

    for n in (2, 3, 10, 103, 1000):
        a, b = O.linspace_f32(1.0, 1e-5, n), torch.linspace(1.0, 1e-5, n).numpy()
        assert np.abs(a - b).max() <= 1.2e-7 and a[0] == b[0] and a[-1] == b[-1]


def test_schedule_closed_form():
    # caching.py:131-181
    assert O.determine_recompute_set(0, 5, 10, 187) == list(range(187))
    assert O.determine_recompute_set(500, 5, 10, 187) == list(range(10))
    assert O.determine_recompute_set(10, 5, 10, 187) == []
    assert O.determine_recompute_set(150, 1, 150, 187) == [0, 1]
    assert O.determine_recompute_set(500, 0, 10, 187) == []
    assert O.determine_recompute_set(500, 200, 10, 187) == list(range(187))


def test_torch_port_matches_reference_vectors(tiny_weights):
    """The timed CPU baseline (oracle/torch_port.py) computes the same thing as the reference."""
    import torch

    from oracle.torch_port import TorchPort, recompute_set

    g = load_golden("tiny_vp_freq_cache.npz")
    port = TorchPort({k: torch.from_numpy(np.array(v)) for k, v in tiny_weights.items()}, 16, 2, 2, "vp", True)
    ts = torch.from_numpy(g["timesteps"])
    dt = ts[0] - ts[1]
    X = port.G[None, :, None] * torch.from_numpy(g["prior_z"][0])
    keep = {int(k): i for i, k in enumerate(g["keep_steps"])}
    with torch.no_grad():
        for i in range(int(g["n_steps"])):
            if i in keep:
                assert rel(X.numpy(), g["keep_X"][keep[i]]) < 2e-4, i
            X = port.step(X, float(ts[i]), torch.from_numpy(g["step_z"][0][i]), dt, recompute_set(i, 2, 100, 16))
    assert port.k_cache is not None and rel(X.numpy(), g["final"][:3]) < 5e-4


FRESCA_KW = {"a": dict(low_scale=1.0, high_scale=1.5, cutoff_ratio=0.5, timestep=0.37, num_steps=1000),
             "b": dict(low_scale=0.8, high_scale=1.2, cutoff_ratio=0.3),
             "c": dict(low_scale=1.3, high_scale=0.9, cutoff_ratio=0.7, timestep=0.9, num_steps=20)}


@pytest.mark.parametrize("L", [187, 100, 24])
@pytest.mark.parametrize("strat", ["spatial", "energy"])
def test_fresca_oracle_matches_reference(L, strat):
    g = load_golden("fresca.npz")
    for tag, kw in FRESCA_KW.items():
        got = O.apply_fresca_to_score(g[f"x_{L}"], cutoff_strategy=strat, **kw)
        assert rel(got, g[f"y_{L}_{strat}_{tag}"]) < 2e-6, (L, strat, tag)


@pytest.mark.parametrize("tag", ["tiny", "ecg", "sat", "nasa"])
def test_event_intensity_oracle_matches_reference(tag):
    """compute_event_intensity (caching.py:524-559): the restatement against values produced by the reference."""
    g = load_golden("event_intensity.npz")
    assert O.event_intensity(None, None, 0.1, step=0) == float(g[f"{tag}_none0"]) == 1.0
    assert O.event_intensity(None, None, 0.1, step=5) == float(g[f"{tag}_none5"]) == 0.1
    got = O.event_intensity(g[f"{tag}_cur"], g[f"{tag}_prev"], float(g[f"{tag}_tau"]), step=3)
    assert abs(got - float(g[f"{tag}_val"])) <= 2e-6 * max(1.0, abs(float(g[f"{tag}_val"])))


def test_emd2_1d_oracle_against_the_transport_lp():
    """The 1-D transport cost behind the Wasserstein metrics (POT's emd2_1d, absent here): the quantile-coupling
    restatement and the literal two-pointer walk against the optimum of the transport linear programme, and
    against closed forms."""
    from scipy.optimize import linprog

    from oracle import metrics_oracle as MO

    rng = np.random.default_rng(3)
    for n, m in [(1, 1), (4, 4), (5, 3), (7, 12), (9, 1), (16, 11)]:
        a, b = rng.normal(size=n), rng.normal(size=m) * 1.7 + 0.3
        C = (a[:, None] - b[None, :]) ** 2
        A_eq = np.zeros((n + m, n * m))
        for i in range(n):
            A_eq[i, i * m:(i + 1) * m] = 1.0
        for j in range(m):
            A_eq[n + j, j::m] = 1.0
        b_eq = np.concatenate([np.full(n, 1.0 / n), np.full(m, 1.0 / m)])
        lp = linprog(C.ravel(), A_eq=A_eq, b_eq=b_eq, bounds=(0, None), method="highs")
        assert lp.status == 0
        assert abs(MO.emd2_1d(a, b) - lp.fun) < 1e-9 * max(1.0, lp.fun)
        assert abs(MO.emd2_1d_walk(a, b) - lp.fun) < 1e-9 * max(1.0, lp.fun)
    x = rng.normal(size=1000)
    assert abs(np.sqrt(MO.emd2_1d(x, x + 2.5)) - 2.5) < 1e-12        # translation
    assert abs(np.sqrt(MO.emd2_1d(x, 3.0 * x)) - 2.0 * np.sqrt(np.mean(x ** 2))) < 1e-12  # scaling: comonotone
    # class level: standardise divides both by std(original)
    X, Y = rng.normal(size=(50, 6)), rng.normal(size=(70, 6)) + 0.5
    d0 = MO.WassersteinDistances(X, Y, seed=1).sliced_distances(5)
    d1 = MO.WassersteinDistances(X, Y, normalisation="standardise", seed=1).sliced_distances(5)
    dirs = MO.WassersteinDistances(X, Y, seed=1).get_random_directions(5)
    assert np.allclose(d1, d0 / np.array([np.std(X @ u) for u in dirs]))


@pytest.mark.parametrize("L", [24, 187, 101, 252])
def test_spectral_density_oracle_matches_reference(L):
    """spectral_density (fourier.py:97-131), odd and even lengths, with and without the inner dft."""
    g = load_golden("spectral.npz")
    x = g[f"x_{L}"].astype(np.float64)
    assert np.allclose(O.spectral_density(x), g[f"sd_{L}"], rtol=1e-4, atol=1e-5)
    assert np.allclose(O.spectral_density(O.dft(x), apply_dft=False), g[f"sdf_{L}"], rtol=1e-4, atol=1e-5)
