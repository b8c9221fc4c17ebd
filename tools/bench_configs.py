"""Times BASELINE.json configs 1, 3, 4 and 5 on one GPU through the drop-in API, inputs resident in
HBM, CUDA events around `reps` calls after warm-up, one JSON line per config with the 8 B/voxel roofline figure.
(cfg 2 is bench.py's default workload; cfg 1 and 3 are also `bench.py --workload cfg1|cfg3`.)  These are records
for profiles/, not the bench contract.  Usage: python tools/bench_configs.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import torch  # noqa: E402

import bench  # noqa: E402
import filters_and_operators as F  # noqa: E402
import stylization_layers as S  # noqa: E402
from mvtb import functional as Fn  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
peak, peak_src = bench.peak_hbm()
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)       # > L2: written between repetitions of the small cases


def timed(fn, voxels, units, unit_name, label, flush_l2):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    gbs = 8.0 * voxels / (ms * 1e-3) / 1e9
    print(json.dumps({"config": label, "ms_per_call_median": round(ms, 4), "value": round(units / (ms * 1e-3), 1),
                      "unit": unit_name + "/s", "algorithmic_GBps_8B_per_voxel": round(gbs, 1),
                      "frac_of_peak": round(gbs / peak, 4), "peak_GBps": peak, "peak_source": peak_src,
                      "l2": "flushed between calls" if flush_l2 else "input + output larger than L2", "reps": reps}), flush=True)


with torch.no_grad():
    # ---- cfg 1: RandFourierDiskMaskd(r=12.5) on 64 x (1,240,240,155), one call per sample batch through the functional API
    cfg = dict(bench.WORKLOADS["cfg1"])
    x = bench.make_inputs(cfg, 0, dev)
    out = torch.empty_like(x)
    timed(lambda: bench.gpu_step(cfg, x, None, out, 0), x.numel(), x.shape[0], "volumes",
          "cfg1: RandFourierDiskMaskd r=12.5, 64 x (1,240,240,155)", False)
    t = F.RandFourierDiskMaskd("image", r=12.5, inside_off=False, prob=1.)
    x1 = x[0]
    timed(lambda: t({"image": x1}), x1.numel(), 1, "volumes",
          "cfg1 (one call of the reference class on one CUDA volume, launch-bound)", True)
    del out

    # ---- cfg 3: disk r=12.5 + S&P(0.15) on 4-channel samples (16 samples = 64 channel volumes per step)
    cfg = dict(bench.WORKLOADS["cfg3"])
    x = bench.make_inputs(cfg, 0, dev)
    out = torch.empty_like(x)
    timed(lambda: bench.gpu_step(cfg, x, None, out, 0), x.numel(), x.shape[0], "samples(4ch)",
          "cfg3: disk r=12.5 + salt-and-pepper 0.15, 16 x (4,240,240,155)", False)
    del x, out

    # ---- cfg 4: 2-D k-space spike on (8192,240,240): same location in every slice (F:982-983)
    x = torch.randn(8192, 240, 240, device=dev)
    t4 = F.KSpaceSpikeNoise((120 + 31, 120 - 17), 15.0)
    timed(lambda: t4(x), x.numel(), x.shape[0], "slices", "cfg4: KSpaceSpikeNoise 2-D, (8192,240,240), one location", False)
    del x

    # ---- cfg 5: (B,1,128,128,64) -> chain-127 -> GibbsNoiseLayer(0.7).forward
    B = 32
    x = torch.randn(B, 1, 128, 128, 64, device=dev)
    idxs = [(64 + (b % 5), 64 - (b % 7), 32 + (b % 3)) for b in range(B)]
    layer = S.GibbsNoiseLayer(0.7)

    def cfg5():
        y = Fn.chain127(x, r=12.5, spike_idx=idxs, intensity=15.0, alpha=0.5, p=0.05, seed=7)
        return layer(y)

    timed(cfg5, 2 * x.numel(), B, "volumes",
          "cfg5: chain-127 then GibbsNoiseLayer(0.7) on (32,1,128,128,64); two transforms = 16 B/voxel", True)
    timed(lambda: layer(x), x.numel(), B, "volumes", "cfg5b: GibbsNoiseLayer(0.7) alone on (32,1,128,128,64)", True)
