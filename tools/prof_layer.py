"""Per-kernel event times of GibbsNoiseLayer(0.7).forward on (B,1,128,128,64) (BASELINE cfg 5's layer), through the
plan's profiling hooks.  Usage: python tools/prof_layer.py [B] [H W D]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import torch  # noqa: E402

import stylization_layers as S  # noqa: E402
from mvtb import _lib, functional as Fn  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
shp = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (128, 128, 64)
dev = torch.device("cuda:0")
x = torch.randn((B, 1) + shp, device=dev)
layer = S.GibbsNoiseLayer(0.7)
L = _lib.lib()
with torch.no_grad():
    for _ in range(3):
        layer(x)
    torch.cuda.synchronize()
    plan = Fn.get_plan((1,) + shp, B, dev)
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    reps = 5
    for _ in range(reps):
        layer(x)
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    tot = 0.0
    for k in range(_lib.K_KINDS):
        if cn[k]:
            print("%-14s launches/call %5.1f  ms/call %.4f  us/volume %.2f" % (L.mvtb_kernel_name(k).decode(), cn[k] / reps, ms[k] / reps, 1e3 * ms[k] / reps / B))
            tot += ms[k] / reps
    print("sum of kernels ms/call %.4f  (%.2f us/volume)" % (tot, 1e3 * tot / B))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        layer(x)
    b.record()
    torch.cuda.synchronize()
    print("wall (events) ms/call %.4f" % (a.elapsed_time(b) / reps))
