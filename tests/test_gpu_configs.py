"""BASELINE.json configs 4 and 5 as parity cases: the 2-D k-space spike on a stack of
240x240 slices, and the on-the-fly chain feeding GibbsNoiseLayer on (B,1,128,128,64) batches."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_cfg4_2d_spike_on_slice_stack(cuda_device):
    """KSpaceSpikeNoise(loc=(X,Y), k_intensity) on (C,240,240): the same location in every slice
    (F:982-983), plus a per-slice-location variant (F:975-979)."""
    import filters_and_operators as F
    from oracle import ref_port as P
    C_ = 24
    x = P.synthetic_volume(40, (C_, 240, 240)) + 0.05
    loc = (120 + 31, 120 - 17)
    y = F.KSpaceSpikeNoise(loc, 12.0)(x.to(cuda_device)).cpu()
    ref = P.kspace_spike(x, loc, 12.0)
    for c in range(C_):
        assert rel_l2(y[c].numpy(), ref[c].numpy()) <= TOL
    locs = tuple((c, 120 + 3 * c - 20, 120 - 2 * c + 9) for c in range(8))
    ints = tuple(10.0 + 0.25 * c for c in range(8))
    x8 = x[:8].contiguous()
    y8 = F.KSpaceSpikeNoise(locs, ints)(x8.to(cuda_device)).cpu()
    ref8 = P.kspace_spike(x8, locs, ints)
    assert rel_l2(y8.numpy(), ref8.numpy()) <= TOL


def test_cfg5_chain_then_gibbs_layer_batch(cuda_device):
    """Per rank: (B,1,128,128,64) -> chain-127 kernels -> GibbsNoiseLayer(0.7).forward (4-D FFT over (1,128,128,64))."""
    import stylization_layers as S
    from mvtb import functional as Fn
    from oracle import ref_port as P
    B_ = 3
    shape = (1, 128, 128, 64)
    xs = [P.synthetic_volume(50 + b, shape) for b in range(B_)]
    us = [P.synthetic_uniform(50 + b, shape) for b in range(B_)]
    idxs = [(64 + 4, 64 - 3, 32 + 2), (64 - 6, 64 + 1, 32 - 5), (64 + 2, 64 + 7, 32 + 3)]      # inside the r = 12.5 ball
    x = torch.stack(xs).to(cuda_device)
    u = torch.stack(us).to(cuda_device)
    y = Fn.chain127(x, r=12.5, spike_idx=idxs, intensity=12.0, alpha=0.5, p=0.05, u=u)
    layer = S.GibbsNoiseLayer(0.7)
    with torch.no_grad():
        z = layer(y)
    assert z.is_cuda and z.shape == x.shape
    for b in range(B_):
        y3 = P.chain_127(xs[b], 12.5, idxs[b], 12.0, 0.5, 0.05, None)
        # S&P is exact given the same input; compare the layer on the GPU's own chain output
        yb = y[b].cpu()
        k3 = Fn.chain127(x[b:b + 1], r=12.5, spike_idx=[idxs[b]], intensity=12.0, alpha=0.5, p=None)[0].cpu()
        assert rel_l2(k3.numpy(), y3.numpy()) <= TOL
        assert torch.equal(yb, P.salt_and_pepper(k3, 0.05, us[b]))
        with torch.no_grad():
            ref = P.gibbs_layer(yb[None], 0.7)[0]
        assert rel_l2(z[b].cpu().numpy(), ref.numpy()) <= TOL


@pytest.mark.parametrize("shape", [(1, 181, 217, 181), (2, 74, 82, 37)])
def test_arbitrary_axis_lengths_prime_factors_above_31(cuda_device, shape):
    """MNI-space 181x217x181 (181 is prime) and other lengths outside 2^a 3^b 5^c ... 31^k: generic direct-DFT stages."""
    import filters_and_operators as F
    from oracle import ref_port as P
    x = P.synthetic_volume(60, shape)
    y = F.GibbsNoise(0.4)(x.to(cuda_device)).cpu()
    assert rel_l2(y.numpy(), P.gibbs_noise(x, 0.4).numpy()) <= TOL
    yd = F.RandFourierDiskMaskd("image", r=12.5, prob=1.)({"image": x.to(cuda_device)})["image"].cpu()
    assert rel_l2(yd.numpy(), P.fourier_disk_mask(x, 12.5, False).numpy()) <= TOL


@pytest.mark.parametrize("shape,alpha", [((2, 240, 240, 155), 0.5), ((1, 240, 240, 155), 0.0), ((3, 64, 48, 31), 0.25)])
def test_wrap_odd_last_axis_dedicated_path(cuda_device, shape, alpha):
    """WrapArtifact on BraTS-shaped volumes (even H, W, odd D): the folds + one-kernel D filter, not the 5-pass chain."""
    import ctypes as C
    import filters_and_operators as F
    from mvtb import _lib, functional as Fn
    from oracle import ref_port as P
    x = P.synthetic_volume(70, shape)
    plan = Fn.get_plan(shape[1:], shape[0], cuda_device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    y = F.WrapArtifact(alpha)(x.to(cuda_device))
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    kinds = {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]}
    assert kinds == {"k_rows_wrap"}
    assert rel_l2(y.cpu().numpy(), P.wrap_artifact(x, alpha).numpy()) <= TOL


def test_cfg4_full_size_spike_property(cuda_device):
    """BASELINE cfg 4 at its full size, (8192, 240, 240): after KSpaceSpikeNoise(loc, I) the bin at loc has magnitude
    e^I and the phase it had, and nothing else in k-space moved (checked with torch.fft on a sample of slices)."""
    import math
    import filters_and_operators as F
    g = torch.Generator(device=cuda_device).manual_seed(4)
    x = torch.randn(8192, 240, 240, generator=g, device=cuda_device)
    loc, inten = (120 + 31, 120 - 17), 9.0
    y = F.KSpaceSpikeNoise(loc, inten)(x)
    assert y.shape == x.shape and y.is_cuda
    pick = torch.tensor([0, 1, 17, 4095, 4096, 8000, 8191], device=cuda_device)
    kx = torch.fft.fftshift(torch.fft.fftn(x[pick].double(), dim=(-2, -1)), dim=(-2, -1))
    ky = torch.fft.fftshift(torch.fft.fftn(y[pick].double(), dim=(-2, -1)), dim=(-2, -1))
    a, b = loc
    a2, b2 = (240 - a) % 240, (240 - b) % 240                  # the conjugate partner of a real image
    want = math.exp(inten) * kx[:, a, b] / kx[:, a, b].abs()
    got = 2 * ky[:, a, b] - kx[:, a, b]                          # taking the real part halves the change at each partner
    assert torch.allclose(got, want, rtol=2e-4, atol=0)
    d = (ky - kx)
    d[:, a, b] = 0
    d[:, a2, b2] = 0
    assert float(d.abs().pow(2).sum().sqrt() / kx.abs().pow(2).sum().sqrt()) <= 1e-5
