"""ncu driver: spike-only chain (RandPlaneWaves_ellipsoid alone) on a few 240x240x155 volumes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
from mvtb import functional as Fn, host  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
cfg = dict(bench.WORKLOADS["cfg2"])
cfg["batch"] = n
x = bench.make_inputs(cfg, 0, dev)
idxs = bench.spike_indices(0, n)
descs = [host.make_desc(spikes=[(idxs[b], host.exp_f32(15.0))]) for b in range(n)]
out = torch.empty_like(x)
for _ in range(2):
    Fn.kspace_chain(x, 3, descs, out=out)
torch.cuda.synchronize()
print("ok", float(out.double().sum()))
