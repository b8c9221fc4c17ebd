"""Small driver for ncu: runs the cfg2 chain (disk r=12.5 + spike + wrap + S&P) on a few 240x240x155
volumes through the public API.  Usage: python tools/prof_chain.py [n_volumes] [repeats] [general]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from mvtb import _lib, functional as Fn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
general = len(sys.argv) > 3 and sys.argv[3] == "general"
cfg = dict(bench.WORKLOADS["cfg2"])
cfg["batch"] = n
dev = torch.device("cuda:0")
x = bench.make_inputs(cfg, 0, dev)
out = torch.empty_like(x)
idxs = bench.spike_indices(0, n)
if general:
    plan = Fn.get_plan(bench.SHAPE, n, dev)
    _lib.check(_lib.lib(), _lib.lib().mvtb_plan_set_path(plan, 1))
for s in range(reps):
    bench.gpu_step(cfg, x, idxs, out, s)
torch.cuda.synchronize()
print("ok", float(out.double().sum()))
