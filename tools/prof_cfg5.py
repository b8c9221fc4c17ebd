"""Where cfg5's time goes: chain-127 and GibbsNoiseLayer(0.7) on (32,1,128,128,64), each timed alone with CUDA events,
with the library's per-kernel profile."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import stylization_layers as S  # noqa: E402
from mvtb import _lib, functional as Fn  # noqa: E402

dev = torch.device("cuda", 0)
L = _lib.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.randn(B, 1, 128, 128, 64, device=dev)
idxs = [(64 + (b % 5), 64 - (b % 7), 32 + (b % 3)) for b in range(B)]
layer = S.GibbsNoiseLayer(float(sys.argv[2]) if len(sys.argv) > 2 else 0.7)
plan = Fn.get_plan((128, 128, 64), B, dev)
plan_layer = Fn.get_plan((1, 128, 128, 64), B, dev)           # the layer transforms over (C, H, W, D)


def timed(name, fn, n=10, plan=plan):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        _lib.check(L, L.mvtb_plan_profile(plan, 1))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
        _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
        _lib.check(L, L.mvtb_plan_profile(plan, 0))
    print(f"{name}: {a.elapsed_time(b) / n * 1000:.1f} us per call of {B} volumes;",
          {L.mvtb_kernel_name(k).decode(): round(ms[k] / n * 1000, 1) for k in range(_lib.K_KINDS) if cn[k]})


y = Fn.chain127(x, r=12.5, spike_idx=idxs, intensity=15.0, alpha=0.5, p=0.05, seed=7, sparse=True)
timed("chain127", lambda: Fn.chain127(x, r=12.5, spike_idx=idxs, intensity=15.0, alpha=0.5, p=0.05, seed=7, sparse=True))
timed("layer", lambda: layer(y), plan=plan_layer)
