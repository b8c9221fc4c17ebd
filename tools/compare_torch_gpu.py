"""Same-box library bar: the reference's device-agnostic routines restated op for op in torch on
the GPU (cuFFT + elementwise kernels, as the reference's own code runs there) next to the mvtb drop-ins.
Measurement tool, not product code.  Usage: python tools/compare_torch_gpu.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import json  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import filters_and_operators as F  # noqa: E402
import stylization_layers as S  # noqa: E402

dev = torch.device("cuda:0")


def kspace(x, n):
    ax = tuple(range(-n, 0))
    return torch.fft.fftshift(torch.fft.fftn(x, dim=ax), dim=ax)


def image(k, n):
    ax = tuple(range(-n, 0))
    return torch.fft.ifftn(torch.fft.ifftshift(k, dim=ax), dim=ax).real


def ref_layer(x, alpha):                      # stylization_layers.py:79-116 on the device
    n = x.dim() - 1
    k = kspace(x, n)
    shape = k.shape[1:]
    centre = (torch.tensor(shape, dtype=torch.float, device=dev) - 1) / 2
    grids = torch.meshgrid([torch.linspace(0, i - 1, i) for i in shape], indexing="ij")
    dist = torch.sqrt(sum((g.to(dev) - c) ** 2 for g, c in zip(grids, centre)))
    an = alpha * dist.max()
    nd = dist / an
    m = nd.where(nd < 1, torch.zeros_like(an))
    m = m.where(nd > 1, torch.ones_like(an))
    m = torch.repeat_interleave(m[None], k.size(0), 0)
    return image(k * m, n)


def ref_gibbs(x, alpha):                      # filters_and_operators.py:663-705 with a CUDA input
    n = x.dim() - 1
    k = kspace(x, n)
    shape = k.shape[1:]
    r = (1 - alpha) * np.max(shape) * np.sqrt(2) / 2.0
    centre = (np.array(shape) - 1) / 2
    grid = np.ogrid[tuple(slice(0, i) for i in shape)]
    mask = np.sqrt(sum((g - c) ** 2 for g, c in zip(grid, centre))) <= r
    mask = np.repeat(mask[None], k.shape[0], axis=0)
    return image(k * torch.tensor(mask, device=k.device), n)


def ref_wrap(x, alpha):                       # filters_and_operators.py:503-515
    k = kspace(x, 3)
    k[:, 1::2, :, :] = k[:, 1::2, :, :] * alpha
    k[:, :, 1::2, :] = k[:, :, 1::2, :] * alpha
    k[:, :, :, 1::2] = k[:, :, :, 1::2] * alpha
    return image(k, 3)


def ref_kspike(x, loc, inten):                # filters_and_operators.py:906-945, one spatial loc, all channels
    k = kspace(x, 3)
    la = torch.log(torch.absolute(k) + 1e-10)
    ph = torch.angle(k)
    la[:, loc[0], loc[1], loc[2]] = inten
    return image(torch.exp(la) * torch.exp(1j * ph), 3)


def timeit(f, reps):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
with torch.no_grad():
    x = torch.randn(2, 1, 128, 128, 64, device=dev)
    alpha = torch.tensor([0.7], device=dev)
    layer = S.GibbsNoiseLayer(0.7)
    rows.append(("GibbsNoiseLayer(0.7) (2,1,128,128,64)", timeit(lambda: ref_layer(x, alpha), 20), timeit(lambda: layer(x), 20)))
    x = torch.randn(4, 128, 128, 64, device=dev)
    g = F.GibbsNoise(0.5)
    rows.append(("GibbsNoise(0.5) (4,128,128,64)", timeit(lambda: ref_gibbs(x, 0.5), 10), timeit(lambda: g(x), 10)))
    w = F.WrapArtifact(0.5)
    rows.append(("WrapArtifact(0.5) (4,128,128,64)", timeit(lambda: ref_wrap(x, 0.5), 10), timeit(lambda: w(x), 10)))
    ks = F.KSpaceSpikeNoise((70, 60, 40), 12.0)
    rows.append(("KSpaceSpikeNoise (4,128,128,64)", timeit(lambda: ref_kspike(x, (70, 60, 40), 12.0), 10), timeit(lambda: ks(x), 10)))
    x = torch.randn(4, 240, 240, 155, device=dev)
    rows.append(("GibbsNoise(0.5) (4,240,240,155)", timeit(lambda: ref_gibbs(x, 0.5), 5), timeit(lambda: g(x), 5)))
    rows.append(("WrapArtifact(0.5) (4,240,240,155)", timeit(lambda: ref_wrap(x, 0.5), 5), timeit(lambda: w(x), 5)))
    ks2 = F.KSpaceSpikeNoise((150, 100, 90), 15.0)
    rows.append(("KSpaceSpikeNoise (4,240,240,155)", timeit(lambda: ref_kspike(x, (150, 100, 90), 15.0), 5), timeit(lambda: ks2(x), 5)))
    d = F.RandFourierDiskMaskd("image", r=12.5, prob=1.)
    k3 = lambda: image(kspace(x, 3) * 1.0, 3)     # noqa: E731  (reference disk mask is CPU-only: FFT part alone)
    rows.append(("RandFourierDiskMaskd(12.5) (4,240,240,155) vs cuFFT fftn+shift+ifftn only", timeit(k3, 5), timeit(lambda: d({"image": x}), 5)))
for name, a, b in rows:
    print(json.dumps({"case": name, "torch_cufft_ms": round(a, 3), "mvtb_ms": round(b, 3), "speedup": round(a / b, 2)}))
