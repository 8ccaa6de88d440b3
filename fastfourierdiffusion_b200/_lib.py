"""ctypes binding of libffd_b200.so (include/ffd_b200.h).

The product path has no CPU fallback: loading fails loudly when the shared
library is missing, and every compute entry point requires CUDA tensors.
PyTorch is used for device memory and streams only.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# FFD_LIB selects another build of the same library (A/B timing of kernel variants, tools/ab.sh) without touching
# the in-tree product file
LIB_PATH = os.environ.get("FFD_LIB") or os.path.join(_PKG, "libffd_b200.so")

PREC_FP32, PREC_FP16 = 0, 1
PREC_BF16 = PREC_FP16  # former name (the tensor-core operands were bf16 until they became fp16)
SDE_VP, SDE_VE = 0, 1
MODE_STOCK, MODE_CACHED = 0, 1

EXPORTS = [
    "ffd_version", "ffd_last_error", "ffd_weights_count", "ffd_layer_stride", "ffd_layer0_offset",
    "ffd_packed_bytes", "ffd_pack_weights", "ffd_workspace_bytes", "ffd_has_tensor_path",
    "ffd_dft_matrix", "ffd_dft_apply", "ffd_posemb_renorm", "ffd_time_embedding", "ffd_score_forward",
    "ffd_sde_step", "ffd_prior", "ffd_philox_normal", "ffd_sample_steps", "ffd_ffn_ln", "ffd_launch_count", "ffd_attn_cached_ln", "ffd_layer_cached_ln", "ffd_fresca", "ffd_fresca_scratch_floats", "ffd_event_intensity",
    "ffd_event_decide", "ffd_event_refresh_kv", "ffd_graph_scratch_bytes",
    "ffd_metric_project", "ffd_metric_sort_rows", "ffd_metric_w2_sorted", "ffd_metric_row_std", "ffd_spectral_density",
]


class FFDError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [("max_len", C.c_int32), ("n_channels", C.c_int32), ("d_model", C.c_int32),
                ("n_head", C.c_int32), ("d_ff", C.c_int32), ("n_layers", C.c_int32)]


class ForwardArgs(C.Structure):
    _fields_ = [("dims", Dims), ("precision", C.c_int32), ("mode", C.c_int32), ("B", C.c_int64),
                ("weights", C.c_void_p), ("packed", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("X", C.c_void_p), ("time_emb", C.c_void_p),
                ("n_own", C.c_int32), ("own_mask", C.c_void_p), ("k_cache", C.c_void_p),
                ("v_cache", C.c_void_p), ("crf_out", C.c_void_p), ("score", C.c_void_p)]


class StepCoef(C.Structure):
    _fields_ = [("kind", C.c_int32), ("c_x", C.c_float), ("coef", C.c_float), ("dt", C.c_float),
                ("sqrt_dt", C.c_float)]


EVENT_STATS = 8


class EventArgs(C.Structure):
    """ffd_event_args: the device-decided E2-CRF schedule (device buffers + thresholds)."""
    _fields_ = [("z_prev", C.c_void_p), ("accum", C.c_void_p), ("weight", C.c_void_p), ("stats", C.c_void_p),
                ("K", C.c_int32), ("R", C.c_int32), ("accumulate", C.c_int32), ("tau_0", C.c_float),
                ("tau_warn", C.c_float), ("eps", C.c_float), ("eta", C.c_float), ("probe_ratio", C.c_float),
                ("seed", C.c_uint64)]


class FrescaArgs(C.Structure):
    """ffd_fresca_args: FreSca inside the fused step loop."""
    _fields_ = [("U", C.c_void_p), ("V", C.c_void_p), ("scratch", C.c_void_p),
                ("high_scale_host", C.POINTER(C.c_float)), ("low_scale", C.c_float), ("cutoff_ratio", C.c_float),
                ("strategy", C.c_int32)]


class SampleArgs(C.Structure):
    _fields_ = [("fwd", ForwardArgs), ("G", C.c_void_p), ("time_emb_table", C.c_void_p),
                ("coef_host", C.POINTER(StepCoef)), ("n_own_host", C.POINTER(C.c_int32)),
                ("store_crf_host", C.POINTER(C.c_int32)), ("noise", C.c_void_p), ("noise_step0", C.c_int32), ("seed", C.c_uint64),
                ("stream_base", C.c_uint64), ("x", C.c_void_p), ("x_tmp", C.c_void_p),
                ("step_begin", C.c_int32), ("step_end", C.c_int32), ("event", C.POINTER(EventArgs)),
                ("fresca", C.POINTER(FrescaArgs)), ("graph", C.c_int32), ("graph_scratch", C.c_void_p),
                ("graph_scratch_bytes", C.c_int64)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the library and declare every prototype of include/ffd_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFDError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    pd = C.POINTER(Dims)
    lib.ffd_version.restype = i32
    lib.ffd_last_error.restype = C.c_char_p
    for n in ("ffd_weights_count", "ffd_layer_stride", "ffd_layer0_offset", "ffd_packed_bytes"):
        getattr(lib, n).restype = i64
        getattr(lib, n).argtypes = [pd]
    lib.ffd_pack_weights.argtypes = [pd, vp, vp, vp]
    lib.ffd_workspace_bytes.restype = i64
    lib.ffd_workspace_bytes.argtypes = [pd, i64, i32]
    lib.ffd_has_tensor_path.restype = i32
    lib.ffd_dft_matrix.argtypes = [vp, i32, i32, vp]
    lib.ffd_dft_apply.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp]
    lib.ffd_posemb_renorm.argtypes = [vp, i32, i32, f32, vp]
    lib.ffd_time_embedding.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp]
    lib.ffd_score_forward.argtypes = [C.POINTER(ForwardArgs), vp]
    lib.ffd_sde_step.argtypes = [C.POINTER(StepCoef), vp, vp, vp, vp, u64, u64, vp, i64, i32, i32, vp]
    lib.ffd_prior.argtypes = [vp, vp, u64, u64, f32, vp, i64, i32, i32, vp]
    lib.ffd_philox_normal.argtypes = [u64, u64, vp, i64, vp]
    lib.ffd_sample_steps.argtypes = [C.POINTER(SampleArgs), vp]
    lib.ffd_ffn_ln.argtypes = [pd, i32, vp, vp, i32, vp, vp, vp, i64, vp]
    lib.ffd_launch_count.restype = u64
    lib.ffd_graph_scratch_bytes.restype = i64
    lib.ffd_graph_scratch_bytes.argtypes = [i32]
    lib.ffd_fresca_scratch_floats.restype = i64
    lib.ffd_fresca_scratch_floats.argtypes = [i64, i32, i32]
    lib.ffd_fresca.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, f32, f32, f32, i32, vp]
    lib.ffd_event_intensity.argtypes = [vp, vp, i64, i32, f32, vp, vp]
    lib.ffd_event_decide.argtypes = [C.POINTER(EventArgs), vp, vp, i32, i32, i32, i32, i32, vp]
    lib.ffd_event_refresh_kv.argtypes = [pd, vp, i32, vp, vp, vp, vp, vp]
    lib.ffd_metric_project.argtypes = [vp, i64, i32, vp, i32, vp, vp]
    lib.ffd_metric_sort_rows.argtypes = [vp, i32, i64, vp]
    lib.ffd_metric_w2_sorted.argtypes = [vp, i64, vp, i64, i32, vp, vp]
    lib.ffd_metric_row_std.argtypes = [vp, i32, i64, vp, vp]
    lib.ffd_spectral_density.argtypes = [vp, vp, i64, i32, i32, vp]
    lib.ffd_attn_cached_ln.argtypes = [pd, i32, vp, vp, i32, vp, vp, vp, vp, vp, i64, i64, vp]
    lib.ffd_layer_cached_ln.argtypes = [pd, i32, vp, vp, i32, vp, vp, vp, vp, vp, i64, i64, vp]
    for n in EXPORTS:
        f = getattr(lib, n)
        if f.restype is C.c_int:  # default: int-returning status
            f.restype = i32
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ffd_last_error().decode("utf-8", "replace")
        raise FFDError(f"ffd_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise FFDError("expected a CUDA tensor: the ffd_b200 path has no CPU fallback")
    if not t.is_contiguous():
        raise FFDError("expected a contiguous tensor")
    return t.data_ptr()


def stream_handle(device: Optional[torch.device] = None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise FFDError("no CUDA device: the ffd_b200 sampling path runs only on a B200 (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())
