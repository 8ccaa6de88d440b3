"""CPU-only checks of the host layer and of the C-ABI library's export table."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

from ffd_testutil import ECG, ROOT, build_seeded_model, load_golden, state_dict_checksums


def _built_lib():
    import __graft_entry__ as g

    g.build()
    from fastfourierdiffusion_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol():
    _lib = _built_lib()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "ffd_b200.h")).read()
    declared = set(re.findall(r"\b(ffd_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().ffd_version() >= 1


def test_struct_layouts_match_header_sizes():
    """sizeof of the ctypes mirrors against sizes compiled from the header with gcc."""
    import subprocess
    import tempfile

    _lib = _built_lib()
    src = '#include "ffd_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(ffd_dims),' \
          'sizeof(ffd_forward_args),sizeof(ffd_step_coef),sizeof(ffd_sample_args),sizeof(ffd_event_args));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                               os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).decode().split()
    got = [ctypes.sizeof(c) for c in (_lib.Dims, _lib.ForwardArgs, _lib.StepCoef, _lib.SampleArgs, _lib.EventArgs)]
    assert got == [int(v) for v in out]


def test_arena_size_and_workspace_query():
    _lib = _built_lib()
    lib = _lib.load()
    d = _lib.Dims(187, 1, 72, 12, 2048, 10)
    n = lib.ffd_weights_count(d)
    # 3 202 413 parameters minus the time encoder (36 + 72*72 + 72), SURVEY 0.7
    assert n == 3202413 - (36 + 72 * 72 + 72)
    assert lib.ffd_layer0_offset(d) + 10 * lib.ffd_layer_stride(d) == n
    assert lib.ffd_workspace_bytes(d, 16384, 0) > 0
    assert lib.ffd_workspace_bytes(d, 0, 0) >= 0


def test_seeded_init_matches_reference_weights():
    """Same constructor order as score_models.py:55-66 -> same weights under the same seed."""
    g = load_golden("ecg_cache.npz")
    m = build_seeded_model(ECG, 42)
    keys, cs = state_dict_checksums(m)
    assert keys == list(g["ck_keys"])
    np.testing.assert_allclose(cs, g["ck_vals"], rtol=1e-12)


def test_state_dict_keys_are_the_reference_keys():
    m = build_seeded_model(ECG, 0, distinct_layers=False)
    keys = set(m.state_dict())
    assert len(keys) == 128
    for k in ("pos_encoder.embedding.weight", "time_encoder.W", "time_encoder.dense.weight", "embedder.weight",
              "unembedder.bias", "backbone.layers.9.self_attn.in_proj_weight", "backbone.layers.0.linear1.weight",
              "backbone.layers.3.norm2.bias"):
        assert k in keys


def test_schedule_and_stats_bookkeeping():
    from fastfourierdiffusion_b200 import E2CRFCache

    c = E2CRFCache(num_layers=10, max_len=187, device=torch.device("cpu"))
    assert c.determine_recompute_set(None, 0.1, 0) == set(range(187))
    assert c.determine_recompute_set(None, 0.1, 500) == set(range(10))
    assert c.determine_recompute_set(None, 0.1, 499) == set()
    c2 = E2CRFCache(num_layers=2, max_len=16, device=torch.device("cpu"), K=2, R=100)
    g = load_golden("tiny_vp_freq_cache.npz")
    for n in g["nS"]:
        c2.account_forward(int(n))
    c2.current_step = 104
    ref = dict(zip(g["stats_keys"], g["stats_vals"]))
    st = c2.get_cache_stats()
    for k in ("recompute_count", "cache_hit_count", "cache_hit_ratio", "cache_ratio", "current_step"):
        assert st[k] == pytest.approx(ref[k], rel=1e-12), k
    with pytest.raises(TypeError):
        E2CRFCache(num_layers=1, max_len=4, device=torch.device("cpu"), random_probe_ratio=0.1)
    with pytest.raises(ZeroDivisionError):
        E2CRFCache(num_layers=1, max_len=4, device=torch.device("cpu"), R=0).crf_stored_at(3)


def test_scheduler_scalars_match_reference_formulas():
    import math

    from fastfourierdiffusion_b200 import VEScheduler, VPScheduler

    vp = VPScheduler(fourier_noise_scaling=True)
    vp.set_noise_scaling(8)
    vp.set_timesteps(10)
    assert vp.G.tolist()[0] == pytest.approx(1.0) and vp.G.tolist()[4] == pytest.approx(1.0)
    assert vp.G.tolist()[1] == pytest.approx(1 / math.sqrt(2))
    c = vp.step_coef(0.5)
    assert c.c_x == pytest.approx(-0.5 * 10.05) and c.coef == pytest.approx(math.sqrt(10.05))
    assert c.dt == pytest.approx(float(vp.timesteps[0] - vp.timesteps[1]))
    ve = VEScheduler(sigma_max=2.0)
    ve.set_noise_scaling(7)
    ve.set_timesteps(5)
    assert ve.G.tolist() == [1.0] * 7
    assert ve.step_coef(1.0).coef == pytest.approx(0.01 * math.sqrt(2 * math.log(200.0)) * 200.0)
    with pytest.raises(IndexError):
        vp.set_timesteps(1)  # sde.py:64


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import fastfourierdiffusion_b200 as F
    from fastfourierdiffusion_b200._lib import FFDError

    with pytest.raises(FFDError):
        F.dft(torch.zeros(1, 8, 1))
    m = build_seeded_model(dict(max_len=8, n_channels=1, d_model=12, num_layers=1, n_head=2), 0)
    b = F.DiffusableBatch(X=torch.zeros(2, 8, 1), timesteps=torch.full((2,), 0.5))
    with pytest.raises(FFDError):
        m(b)
    with pytest.raises(FFDError):
        F.DiffusionSampler(m, 2).sample(2, 4)


def test_event_driven_cache_kwargs_and_schedule():
    """event_driven is an extra, default-off constructor argument: the index schedule is untouched without it, only
    the full step stays host-scheduled with it, and R must be positive."""
    import pytest
    import torch

    from fastfourierdiffusion_b200.caching import E2CRFCache

    c = E2CRFCache(num_layers=2, max_len=16, device=torch.device("cpu"), K=3, R=10)
    assert not c.event_driven and c.determine_recompute_set(None, 0.1, 500) == set(range(6))
    e = E2CRFCache(num_layers=2, max_len=16, device=torch.device("cpu"), K=3, R=10, event_driven=True,
                   event_probe_ratio=0.1, event_seed=5)
    assert e.determine_recompute_set(None, 0.1, 0) == set(range(16))
    assert e.determine_recompute_set(None, 0.1, 500) == set() and e.determine_recompute_set(None, 0.1, 7) == set()
    assert e.get_event_stats() == {}
    with pytest.raises(ValueError):
        E2CRFCache(num_layers=2, max_len=16, device=torch.device("cpu"), R=0, event_driven=True)
    with pytest.raises(TypeError):
        E2CRFCache(num_layers=2, max_len=16, device=torch.device("cpu"), not_an_argument=1)
