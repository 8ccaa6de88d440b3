"""Build libffd_b200.so in-tree with nvcc for sm_100a (no torch headers involved)."""

from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libffd_b200.so")
SOURCES = ["ffd_api.cu", "ffd_simt.cu", "ffd_metrics.cu", "ffd_event.cu", "ffd_tc.cu", "ffd_tc_attn.cu", "ffd_tc_layer.cu"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ffd_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, out: str = "") -> str:
    """``out`` (or FFD_LIB_OUT): write a VARIANT build somewhere else (A/B timing, stamp builds) instead of the product
    library; select it at run time with FFD_LIB."""
    out = out or os.environ.get("FFD_LIB_OUT", "")
    if not out and not force and not needs_build():
        return LIB
    # The CUDA runtime is linked statically on purpose: the library is loaded next to torch, whose wheel brings its own
    # libcudart.so.12 of a different minor version (12.8 vs this toolkit's 12.9); a shared runtime would bind our
    # 12.9-compiled host stubs to whichever copy the process loaded first.
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    cmd += os.environ.get("FFD_NVCC_FLAGS", "").split()
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out or LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out or LIB


if __name__ == "__main__":
    import sys

    print(build(force=True, verbose="-v" in sys.argv))
