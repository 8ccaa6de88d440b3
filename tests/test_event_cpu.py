"""The event-driven schedule of the reference's paper (section 3.3, Eq. 14-18, Algorithm 1) as restated in
oracle/fdiff_oracle.py, against closed forms worked by hand.  The reference's code never runs this schedule, so
there are no reference vectors for it (parity unpinned by reference runs; the GPU tests compare the device path
with this restatement)."""

import numpy as np

from oracle import fdiff_oracle as O


def _state(L=6, d=4):
    st = O.EventState(L, d, np.dtype(np.float64))
    z0 = np.arange(L * d, dtype=np.float64).reshape(L, d) / 10.0 + 1.0
    O.event_decide(O.EventConfig(), st, z0, np.ones((L, 2)), 0, True)
    return st, z0


def test_first_call_initialises():
    st, z0 = _state()
    assert np.array_equal(st.z_prev, z0)
    assert not st.accum.any() and not st.weight.any() and not st.stats.any()


def test_eq14_eq15_closed_form():
    L, d = 6, 4
    st, z0 = _state(L, d)
    cfg = O.EventConfig(K=1, R=1000, tau_0=0.3, tau_warn=10.0, eps=0.0, eta=0.0)
    z1 = z0.copy()
    z1[3] += 0.2        # delta_3 = 0.2 * sqrt(4) = 0.4
    z1[4] += 0.05       # delta_4 = 0.1
    x0 = np.ones((L, 2))
    x0[4] = [2.0, 1.0]  # energy 5 -> tau_4 = 0.06 ; others energy 2 -> tau = 0.15
    O.event_decide(cfg, st, z1, x0, 1, False)
    want_r = (4 * 0.2 ** 2 + 4 * 0.05 ** 2) / (z0 ** 2).sum()
    assert np.isclose(st.last["r"], want_r, rtol=1e-12)
    assert np.allclose(st.last["delta"], [0, 0, 0, 0.4, 0.1, 0], atol=1e-12)
    assert np.allclose(st.last["tau"], [0.15, 0.15, 0.15, 0.15, 0.06, 0.15])
    # tokens 0..K always, 3 and 4 over their thresholds
    assert st.last["rec"].tolist() == [True, True, False, True, True, False]
    assert st.weight.tolist() == [1, 1, 0, 1, 1, 0]
    assert st.stats[1] == 4 and st.stats[2] == 0 and st.stats[3] == 0


def test_error_feedback_accumulator_triggers_slow_drift():
    L, d = 6, 4
    cfg = O.EventConfig(K=0, R=1000, tau_0=0.3, tau_warn=10.0, eps=0.0, eta=0.0)  # tau = 0.15
    for accumulate, want_step in ((True, 4), (False, None)):
        st, z = _state(L, d)
        cfg.accumulate = accumulate
        fired = None
        for i in range(1, 8):
            z = z.copy()
            z[2] += 0.02  # delta = 0.04 per step: below tau, 4 steps add up to 0.16 > 0.15
            O.event_decide(cfg, st, z, np.ones((L, 2)), i, False)
            if st.weight[2] == 1 and fired is None:
                fired = i
                assert st.accum[2] == 0  # reset on recompute
        assert fired == want_step


def test_feedback_step_gamma_and_period():
    L, d = 6, 4
    st, z0 = _state(L, d)
    cfg = O.EventConfig(K=0, R=3, tau_0=1e9, tau_warn=10.0)
    z = z0 * 1.1  # r = 0.01
    O.event_decide(cfg, st, z, np.ones((L, 2)), 1, False)   # (1 + 1) % 3 != 0
    assert st.weight.tolist() == [1, 0, 0, 0, 0, 0]
    O.event_decide(cfg, st, z * 1.1, np.ones((L, 2)), 2, False)  # (2 + 1) % 3 == 0: feedback step
    r = st.last["r"]
    assert np.isclose(r, 0.01, rtol=1e-4)
    assert np.allclose(st.weight[1:], min(0.1, r / 2)) and st.weight[0] == 1
    assert st.stats[3] == 1 and st.stats[2] == L - 1
    cfg2 = O.EventConfig(K=0, R=1000, tau_0=1e9, tau_warn=0.005)
    O.event_decide(cfg2, st, z * 1.21 * 1.1, np.ones((L, 2)), 3, False)  # r > tau_warn -> feedback
    assert bool(st.last["feedback"]) and np.allclose(st.weight[1:], min(0.1, st.last["r"] / 2))


def test_random_probe_is_deterministic_and_sparse():
    u = O.event_uniform(7, 11, 4096)
    assert np.array_equal(u, O.event_uniform(7, 11, 4096))
    assert not np.array_equal(u, O.event_uniform(7, 12, 4096))
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02
    L, d = 256, 4
    st = O.EventState(L, d, np.dtype(np.float64))
    z0 = np.ones((L, d))
    O.event_decide(O.EventConfig(), st, z0, np.ones((L, 1)), 0, True)
    cfg = O.EventConfig(K=3, R=1000, tau_0=1e9, tau_warn=10.0, probe_ratio=0.1, seed=7)
    O.event_decide(cfg, st, z0, np.ones((L, 1)), 5, False)
    probes = st.last["probe"]
    assert not probes[:4].any() and 5 < probes.sum() < 60
    assert np.array_equal(st.weight == 1, probes | (np.arange(L) <= 3))


def test_refresh_blend_and_batch1_equals_mixed_branch():
    """weight 1 on a token set S = the reference's mixed branch for batch size 1 (cached_transformer.py:259-305):
    same K/V rows stored, same layer output."""
    rng = np.random.default_rng(0)
    L, d, H = 10, 8, 2
    lw = O.LayerWeights(**{f: rng.normal(size=s).astype(np.float64) * 0.3 for f, s in dict(
        in_proj_w=(3 * d, d), in_proj_b=(3 * d,), out_w=(d, d), out_b=(d,), w1=(16, d), b1=(16,), w2=(d, 16),
        b2=(d,), n1w=(d,), n1b=(d,), n2w=(d,), n2b=(d,)).items()})
    x = rng.normal(size=(1, L, d))
    def fresh_store():
        st = O.KVStore(1, L)
        st.set_batch(0, range(L), rng.normal(size=(1, H, L, d // H)), rng.normal(size=(1, H, L, d // H)))
        return st
    rng = np.random.default_rng(1); a = fresh_store()
    rng = np.random.default_rng(1); b = fresh_store()
    S = [0, 1, 4, 7]
    want = O.cached_layer(lw, x, H, S, a, 0)
    w = np.zeros(L); w[S] = 1
    O.event_refresh_kv(lw, x[0], w, b, 0, H)
    got = O.cached_layer(lw, x, H, [], b, 0)
    assert np.allclose(a.k, b.k) and np.allclose(a.v, b.v)
    assert np.allclose(got, want, atol=1e-12)
    # partial weight: cache moves a fraction gamma towards the recomputed rows, other rows untouched
    k_before = b.k.copy()
    x2 = rng.normal(size=(1, L, d))
    w2 = np.zeros(L); w2[3] = 0.25
    O.event_refresh_kv(lw, x2[0], w2, b, 0, H)
    kn = O._split_heads((x2[0] @ lw.in_proj_w[d:2 * d].T + lw.in_proj_b[d:2 * d])[None], H)[0]
    assert np.allclose(b.k[0][:, 3], k_before[0][:, 3] + 0.25 * (kn[:, 3] - k_before[0][:, 3]))
    keep = [i for i in range(L) if i != 3]
    assert np.array_equal(b.k[0][:, keep], k_before[0][:, keep])
