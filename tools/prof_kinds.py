"""Per-kernel-kind event times (the plan's profiling hooks) of one transform on a batch.
Usage: python tools/prof_kinds.py spike|gibbs|wrap|layer [B] [H W D]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import torch  # noqa: E402

import filters_and_operators as F  # noqa: E402
import stylization_layers as S  # noqa: E402
from mvtb import _lib, functional as Fn  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "spike"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
shp = tuple(int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (240, 240, 155)
dev = torch.device("cuda:0")
x = torch.randn((B,) + shp, device=dev)
if what == "spike":
    t = F.KSpaceSpikeNoise((shp[0] // 2 + 31, shp[1] // 2 - 17, shp[2] // 2 + 5), 15.0)
elif what == "gibbs":
    t = F.GibbsNoise(0.5)
elif what == "wrap":
    t = F.WrapArtifact(0.5)
else:
    t = S.GibbsNoiseLayer(0.7)
    x = x[:, None]
L = _lib.lib()
with torch.no_grad():
    for _ in range(3):
        t(x)
    torch.cuda.synchronize()
    plan = Fn.get_plan(((1,) + shp) if what == "layer" else shp, B, dev)
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    reps = 5
    for _ in range(reps):
        t(x)
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    tot = 0.0
    for k in range(_lib.K_KINDS):
        if cn[k]:
            print("%-16s launches/call %5.1f  ms/call %.4f  us/volume %.2f" % (L.mvtb_kernel_name(k).decode(), cn[k] / reps, ms[k] / reps, 1e3 * ms[k] / reps / B))
            tot += ms[k] / reps
    print("sum of kernels ms/call %.4f  (%.2f us/volume)" % (tot, 1e3 * tot / B))
