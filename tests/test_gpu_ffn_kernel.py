"""The fused FFN + residual + LayerNorm kernel on its own (ffd_ffn_ln), fp32 CUDA-core path against
a float64 restatement, tcgen05 bf16 path against the fp32 path, ragged row counts included."""

import numpy as np
import pytest
import torch

from ffd_testutil import ECG, TINY, build_seeded_model, build_tiny_model, rel

pytestmark = pytest.mark.gpu


@torch.no_grad()
def _ffn_ref(model, layer, x):
    lyr = model.backbone.layers[layer]
    w1, b1 = lyr.linear1.weight.double(), lyr.linear1.bias.double()
    w2, b2 = lyr.linear2.weight.double(), lyr.linear2.bias.double()
    xd = x.double()
    h = torch.relu(xd @ w1.T + b1) @ w2.T + b2 + xd
    return torch.nn.functional.layer_norm(h, (x.shape[1],), lyr.norm2.weight.double(), lyr.norm2.bias.double(), 1e-5)


def _run(model, layer, x, precision):
    from fastfourierdiffusion_b200 import _lib

    model.set_precision(precision)
    arena, packed = model.device_weights()
    y = torch.empty_like(x)
    scratch = torch.empty_like(x)
    _lib.check(_lib.load().ffd_ffn_ln(model.dims, 1 if precision == "fp16" else 0, _lib.ptr(arena), _lib.ptr(packed),
                                      layer, _lib.ptr(x), _lib.ptr(y), _lib.ptr(scratch), x.shape[0],
                                      _lib.stream_handle()))
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("which", ["tiny", "ecg"])
@pytest.mark.parametrize("M", [1, 100, 256, 300, 5000, 187 * 64])
def test_ffn_ln_fp32_and_tensor_path(which, M):
    from fastfourierdiffusion_b200 import _lib

    model = build_tiny_model() if which == "tiny" else build_seeded_model(ECG, 42).to("cuda").eval()
    d = model.d_model
    layer = len(model.backbone.layers) - 1
    torch.manual_seed(M)
    x = torch.randn(M, d, device="cuda") * 1.3
    want = _ffn_ref(model, layer, x)
    got32 = _run(model, layer, x, "fp32")
    assert rel(got32.cpu().numpy(), want.cpu().numpy()) < 5e-6
    if not _lib.load().ffd_has_tensor_path() or not model._tc_ok():
        pytest.skip("tensor path not available")
    got16 = _run(model, layer, x, "fp16")
    assert torch.isfinite(got16).all()
    # fp16 operands (11-bit mantissa), fp32 accumulate, fp32 residual and LayerNorm
    e_max = rel(got16.cpu().numpy(), want.cpu().numpy())
    err = (got16.double() - want).abs()
    e_mean = float(err.mean() / want.abs().mean())
    print(f"ffn_ln [{which}, M={M}] fp16 operands vs float64: max-rel {e_max:.2e}, mean-rel {e_mean:.2e}")
    assert e_max < 2e-3 and e_mean < 3e-4
    # in place
    xin = x.clone()
    arena, packed = model.device_weights()
    _lib.check(_lib.load().ffd_ffn_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), layer, _lib.ptr(xin),
                                      _lib.ptr(xin), None, M, _lib.stream_handle()))
    assert torch.equal(xin, got16)


@pytest.mark.parametrize("which", ["tiny", "ecg"])
@pytest.mark.parametrize("B", [1, 3, 70])
def test_attn_cached_ln_fp32_and_tensor_path(which, B):
    """The pure-cache attention block on its own (ffd_attn_cached_ln): fp32 path against a float64
    restatement of cached_transformer.py:228-258,309-321, tcgen05 path against the same."""
    from fastfourierdiffusion_b200 import _lib

    model = build_tiny_model() if which == "tiny" else build_seeded_model(ECG, 42).to("cuda").eval()
    d, L, H = model.d_model, model.max_len, model.n_head
    hd = d // H
    NL = len(model.backbone.layers)
    layer = NL - 1
    torch.manual_seed(B)
    x = torch.randn(B * L, d, device="cuda") * 1.5
    kc = torch.randn(NL, H, L, hd, device="cuda")
    vc = torch.randn(NL, H, L, hd, device="cuda")
    lyr = model.backbone.layers[layer]
    with torch.no_grad():
        xd = x.double().view(B, L, d)
        q = (xd @ lyr.self_attn.in_proj_weight[:d].double().T + lyr.self_attn.in_proj_bias[:d].double())
        q = q.view(B, L, H, hd).transpose(1, 2)
        s = q @ kc[layer].double().transpose(-2, -1) / hd ** 0.5
        o = (torch.softmax(s, -1) @ vc[layer].double()).transpose(1, 2).reshape(B, L, d)
        o = o @ lyr.self_attn.out_proj.weight.double().T + lyr.self_attn.out_proj.bias.double()
        want = torch.nn.functional.layer_norm(xd + o, (d,), lyr.norm1.weight.double(), lyr.norm1.bias.double(),
                                              1e-5).view(B * L, d)

    def run(prec):
        model.set_precision(prec)
        arena, packed = model.device_weights()
        ws = model.workspace(B)
        y = torch.empty_like(x)
        _lib.check(_lib.load().ffd_attn_cached_ln(model.dims, 1 if prec == "fp16" else 0, _lib.ptr(arena),
                                                  _lib.ptr(packed), layer, _lib.ptr(kc), _lib.ptr(vc), _lib.ptr(x),
                                                  _lib.ptr(y), _lib.ptr(ws), ws.numel(), B * L, _lib.stream_handle()))
        torch.cuda.synchronize()
        return y

    assert rel(run("fp32").cpu().numpy(), want.cpu().numpy()) < 5e-6
    if not _lib.load().ffd_has_tensor_path() or not model._tc_ok():
        pytest.skip("tensor path not available")
    got = run("fp16")
    assert torch.isfinite(got).all()
    e_max = rel(got.cpu().numpy(), want.cpu().numpy())
    e_mean = float((got.double() - want).abs().mean() / want.abs().mean())
    print(f"attn_cached_ln [{which}, B={B}] fp16 operands vs float64: max-rel {e_max:.2e}, mean-rel {e_mean:.2e}")
    assert e_max < 4e-3 and e_mean < 6e-4
    # large scores (|s| ~ 1e2): the split-K softmax must stay exact where a bounded shift would underflow
    kc2 = kc * 40
    model.set_precision("fp16")
    arena, packed = model.device_weights()
    ws = model.workspace(B)
    y2 = torch.empty_like(x)
    _lib.check(_lib.load().ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), layer, _lib.ptr(kc2),
                                              _lib.ptr(vc), _lib.ptr(x), _lib.ptr(y2), _lib.ptr(ws), ws.numel(), B * L,
                                              _lib.stream_handle()))
    with torch.no_grad():
        s2 = q @ kc2[layer].double().transpose(-2, -1) / hd ** 0.5
        o2 = (torch.softmax(s2, -1) @ vc[layer].double()).transpose(1, 2).reshape(B, L, d)
        o2 = o2 @ lyr.self_attn.out_proj.weight.double().T + lyr.self_attn.out_proj.bias.double()
        want2 = torch.nn.functional.layer_norm(xd + o2, (d,), lyr.norm1.weight.double(), lyr.norm1.bias.double(),
                                               1e-5).view(B * L, d)
    assert torch.isfinite(y2).all()
    assert float((y2.double() - want2).abs().mean() / want2.abs().mean()) < 3e-2


@pytest.mark.parametrize("M", [1, 129, 511, 513, 148 * 256 - 5, 148 * 256 * 2 + 300, 148 * 256 * 5 + 77])
def test_ffn_cta_pair_kernel_equals_one_cta_kernel(M):
    """The cta_group::2 kernel (default for d_model 72) against the one-CTA kernel, same inputs: bit-identical for
    every row count -- ragged last tiles, clusters with an empty second CTA, several tile-pair iterations per
    cluster (a miscompiled X staging once dropped the bias-low column from the second iteration on)."""
    from fastfourierdiffusion_b200 import _lib

    lib = _lib.load()
    if not lib.ffd_has_tensor_path():
        pytest.skip("tensor path not available")
    model = build_seeded_model(ECG, 42, precision="fp16").to("cuda").eval()
    torch.manual_seed(M)
    x = torch.randn(M, model.d_model, device="cuda") * 1.3
    try:
        lib.ffd_debug_set_ffn_pairs(0)
        one = _run(model, 3, x, "fp16")
        lib.ffd_debug_set_ffn_pairs(1)
        pair = _run(model, 3, x, "fp16")
        # identical tiles at different iterations of the persistent loop give identical rows
        if M >= 148 * 256 * 2:
            xx = torch.cat([x[:148 * 256], x[:148 * 256], x[:M - 2 * 148 * 256]]).contiguous()
            yy = _run(model, 3, xx, "fp16")
            assert torch.equal(yy[:148 * 256], yy[148 * 256:2 * 148 * 256])
    finally:
        lib.ffd_debug_set_ffn_pairs(-1)
    assert torch.equal(one, pair)
    assert rel(pair.cpu().numpy(), _ffn_ref(model, 3, x).cpu().numpy()) < 1e-2


@pytest.mark.parametrize("dims", [TINY, dict(max_len=24, n_channels=2, d_model=60, num_layers=2, n_head=12),
                                  dict(max_len=24, n_channels=2, d_model=40, num_layers=1, n_head=8), ECG])
def test_ffn_rows_do_not_depend_on_the_loop_iteration(dims):
    """The persistent FFN kernels (every d_model / DP instantiation) give bit-identical rows for identical tiles
    whichever iteration of a CTA's tile-pair loop processes them."""
    from fastfourierdiffusion_b200 import _lib

    if not _lib.load().ffd_has_tensor_path():
        pytest.skip("tensor path not available")
    model = build_seeded_model(dims, 7, precision="fp16").to("cuda").eval()
    if not model._tc_ok():
        pytest.skip("model shape not supported by the tcgen05 kernels")
    blk = 148 * 256  # rows one wave of CTAs (two tiles each) covers
    torch.manual_seed(3)
    x0 = torch.randn(blk, model.d_model, device="cuda")
    x = torch.cat([x0, x0, x0, x0[:1000]]).contiguous()
    y = _run(model, len(model.backbone.layers) - 1, x, "fp16")
    assert torch.equal(y[:blk], y[blk:2 * blk]) and torch.equal(y[:blk], y[2 * blk:3 * blk])
    assert torch.equal(y[:1000], y[3 * blk:])
    assert rel(y[:blk].cpu().numpy(), _ffn_ref(model, len(model.backbone.layers) - 1, x0).cpu().numpy()) < 1e-2


def test_attention_rows_do_not_depend_on_the_loop_iteration():
    """Same property for the pure-cache attention block (rows are independent there: the K/V are the shared cache):
    identical 128-row tiles handled by different iterations of a CTA's tile loop give bit-identical rows."""
    from fastfourierdiffusion_b200 import _lib

    lib = _lib.load()
    if not lib.ffd_has_tensor_path():
        pytest.skip("tensor path not available")
    model = build_seeded_model(ECG, 42, precision="fp16").to("cuda").eval()
    L, d, H = ECG["max_len"], ECG["d_model"], ECG["n_head"]
    blk = 148 * 128
    M = (3 * blk + L - 1) // L * L
    torch.manual_seed(5)
    x0 = torch.randn(blk, d, device="cuda")
    x = torch.cat([x0, x0, x0, x0[:M - 3 * blk]]).contiguous()
    y = torch.empty_like(x)
    kc = torch.randn(ECG["num_layers"], H, L, d // H, device="cuda")
    vc = torch.randn_like(kc)
    arena, packed = model.device_weights()
    ws = model.workspace(M // L)
    _lib.check(lib.ffd_attn_cached_ln(model.dims, 1, _lib.ptr(arena), _lib.ptr(packed), 2, _lib.ptr(kc), _lib.ptr(vc),
                                      _lib.ptr(x), _lib.ptr(y), _lib.ptr(ws), ws.numel() * ws.element_size(), M,
                                      _lib.stream_handle()))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    assert torch.equal(y[:blk], y[blk:2 * blk]) and torch.equal(y[:blk], y[2 * blk:3 * blk])
    assert torch.equal(y[:M - 3 * blk], y[3 * blk:])
