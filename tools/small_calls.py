"""Host-to-completion time of the small calls (one 240x240x155 volume through RandFourierDiskMaskd, two 128x128x64
volumes through GibbsNoiseLayer): wall clock around call + synchronize, median of 200, and a cProfile of the first."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import filters_and_operators as F  # noqa: E402
import stylization_layers as S  # noqa: E402

dev = torch.device("cuda", 0)
x1 = {"image": torch.randn(1, 240, 240, 155, device=dev)}
t1 = F.RandFourierDiskMaskd(keys="image", r=12.5, inside_off=False, prob=1.0)
x2 = torch.randn(2, 1, 128, 128, 64, device=dev)
layer = S.GibbsNoiseLayer(0.7)


def wall(fn, n=200):
    ts = []
    with torch.no_grad():
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e6


print(f"RandFourierDiskMaskd, 1 x 240x240x155: {wall(lambda: t1(dict(x1))):.1f} us host-to-completion")
print(f"GibbsNoiseLayer(0.7), 2 x 128x128x64:  {wall(lambda: layer(x2)):.1f} us host-to-completion")
pr = cProfile.Profile()
with torch.no_grad():
    pr.enable()
    for _ in range(200):
        t1(dict(x1))
    torch.cuda.synchronize()
    pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
