"""Builds and loads the DEBUG emulator build of the kernel sources (tests only).

g++ -DMVTB_EMU compiles medical-vision-textural-bias_b200/mvtb/csrc/*.cu against
tests/cuemu/cuemu.h into tests/cuemu/_build/libmvtb_emu.so.  It exists so that kernel index
arithmetic can be exercised in a container without a GPU; it is not a product path and
nothing under medical-vision-textural-bias_b200/ can load it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "medical-vision-textural-bias_b200", "mvtb", "csrc")
OUT = os.path.join(HERE, "_build", "libmvtb_emu.so")
SOURCES = ["plan.cu", "kspace_chain.cu", "voxel_ops.cu", "bandlimited.cu", "spike_fast.cu", "intensity.cu", "dice.cu", "spatial.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))] + [os.path.join(HERE, "cuemu.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(HERE, "cuemu.h"),
                                                                                            os.path.join(ROOT, "include", "mvtb.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-shared", "-fPIC", "-DMVTB_EMU", "-I" + HERE, "-x", "c++"] + srcs + ["-o", OUT]
    subprocess.run(cmd, check=True, capture_output=True)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        import sys
        sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
        from mvtb import _lib as B
        _lib = B.bind(C.CDLL(build()))
    return _lib


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Plan:
    def __init__(self, fft_shape, chunk=2):
        from mvtb import _lib as B
        self.h = C.c_void_p()
        shp = (C.c_int * len(fft_shape))(*fft_shape)
        B.check(lib(), lib().mvtb_plan_create(C.byref(self.h), len(fft_shape), shp, chunk, 0))
        self.fft_shape = tuple(fft_shape)

    def __del__(self):
        if getattr(self, "h", None):
            lib().mvtb_plan_destroy(self.h)
            self.h = None


def chain(x, ndim_fft, descs, chunk=2, minmax_vols_per_sample=None):
    """x: float32 numpy array; FFT over the last ndim_fft axes; returns (y, minmax or None)."""
    from mvtb import _lib as B, host
    x = np.ascontiguousarray(x, dtype=np.float32)
    plan = Plan(x.shape[-ndim_fft:], chunk)
    nvol = int(np.prod(x.shape[:-ndim_fft])) if x.ndim > ndim_fft else 1
    y = np.empty_like(x)
    arr = host.desc_array(descs)
    mm = None
    vps = 1
    if minmax_vols_per_sample:
        vps = minmax_vols_per_sample
        mm = np.zeros(2 * ((nvol + vps - 1) // vps), dtype=np.float32)
    B.check(lib(), lib().mvtb_kspace_chain_f32(plan.h, ptr(x), ptr(y), nvol, arr, len(descs), ptr(mm), vps, None))
    return y, mm
