"""Band-limited (pruned-DFT) path on the GPU: same numbers as the general FFT path and the oracle.
Both paths are reached through the same C-ABI call; mvtb_plan_set_path forces the general one."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5


def chain(x, descs, general, **kw):
    """general: False/0 automatic, True/1 general FFT path, 2 pair-folding H kernels, 3 split W/D stage (mvtb.h)."""
    from mvtb import _lib, functional as Fn
    nvol = int(np.prod(x.shape[:-3]))
    plan = Fn.get_plan(tuple(x.shape[-3:]), nvol, x.device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_set_path(plan, int(general)))
    try:
        return Fn.kspace_chain(x, 3, descs, **kw)
    finally:
        _lib.check(L, L.mvtb_plan_set_path(plan, 0))


def disk(thr, **kw):
    from mvtb import _lib, host
    return host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=thr, **kw)


@pytest.mark.parametrize("shape,r", [((1, 240, 240, 155), 12.5), ((2, 128, 128, 64), 12.5), ((1, 128, 128, 64), 9.0),
                                     ((1, 240, 240, 155), 15.0), ((4, 31, 45, 27), 3.5), ((1, 64, 48, 155), 2.0),
                                     ((1, 240, 240, 155), 20.0), ((1, 128, 128, 64), 25.0), ((1, 240, 240, 155), 30.0)])
def test_bl_vs_oracle_and_general(cuda_device, shape, r):
    from mvtb import host
    from oracle import ref_port as P
    x = P.synthetic_volume(11, shape)
    d = disk(host.disk_threshold(r, shape[-3:]))
    yb = chain(x.to(cuda_device), [d], False).cpu().numpy()
    yg = chain(x.to(cuda_device), [d], True).cpu().numpy()
    ref = P.fourier_disk_mask(x, r, False).numpy()
    assert rel_l2(yb, ref) <= TOL and rel_l2(yg, ref) <= TOL and rel_l2(yb, yg) <= TOL
    for variant in (2, 3):                        # pair-folding H kernels; W/D stage as three kernels
        yv = chain(x.to(cuda_device), [d], variant).cpu().numpy()
        assert rel_l2(yv, ref) <= TOL


@pytest.mark.parametrize("split", [False, True])
def test_bl_kernels_are_the_ones_launched(cuda_device, split):
    import ctypes as C
    from mvtb import _lib, functional as Fn, host
    x = torch.randn(2, 128, 128, 64, device=cuda_device)
    plan = Fn.get_plan((128, 128, 64), 2, cuda_device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_set_path(plan, 3 if split else 0))
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    Fn.kspace_chain(x, 3, [disk(host.disk_threshold(12.5, (128, 128, 64)))])
    torch.cuda.synchronize()
    _lib.check(L, L.mvtb_plan_set_path(plan, 0))
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    kinds = {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]}
    # the forward H pass runs on the tensor cores by default (128 % 16 == 0): k_bl_fwd_tc
    assert kinds == {"k_bl_fwd_tc", "k_bl_mid", "k_bl_inv_h"} | ({"k_bl_fwd_w", "k_bl_inv_w"} if split else set())


def test_bl_chain127_full_size_out_of_ball_spike(cuda_device):
    """The 127-chain at 240x240x155 with the spike on the (55,55,30) shell, i.e. on a bin the disk zeroed.
    Both paths define angle(0) = 0 there and must agree; against the reference (whose phase is rounding
    noise) the comparison is modulo that plane wave, with its amplitude checked."""
    from mvtb import host
    from oracle import ref_port as P
    from test_gpu_parity import _remove_plane_wave
    shape = (1, 240, 240, 155)
    x = P.synthetic_volume(0, shape)
    shell = host.ellipsoid_shell(shape[1:], 55., 55., 30.)
    idx = tuple(int(v) for v in shell[np.random.RandomState(0).randint(0, len(shell))])
    d = disk(host.disk_threshold(12.5, shape[1:]), spikes=[(idx, host.exp_f32(15.0))], wrap_alpha=0.5)
    yb, mm = chain(x.to(cuda_device), [d], False, want_minmax=True)
    yg = chain(x.to(cuda_device), [d], True)
    assert rel_l2(yb.cpu().numpy(), yg.cpu().numpy()) <= TOL
    assert float(mm[0, 0]) == float(yb.min()) and float(mm[0, 1]) == float(yb.max())
    ref = P.chain_127(x, 12.5, idx, 15.0, 0.5, 0.0, None).numpy()
    ra, aa = _remove_plane_wave(yb.cpu().numpy(), idx)
    rb, ab = _remove_plane_wave(ref, idx)
    assert rel_l2(ra, rb) <= 1e-4
    assert np.allclose(aa, ab, rtol=1e-3)


def test_bl_batch_per_sample_spikes_and_fused_minmax(cuda_device):
    from mvtb import functional as Fn, host
    from oracle import ref_port as P
    B_ = 5
    x = torch.stack([P.synthetic_volume(i, (1, 128, 128, 64)) for i in range(B_)]).to(cuda_device)
    shell = host.ellipsoid_shell((128, 128, 64), 55., 55., 30.)
    idxs = [tuple(int(v) for v in shell[np.random.RandomState(i).randint(0, len(shell))]) for i in range(B_)]
    idxs[2] = (64 + 3, 64 - 2, 32 + 1)          # one sample with its spike inside the ball
    y = Fn.chain127(x, r=12.5, spike_idx=idxs, intensity=15.0, alpha=0.5, p=None)
    thr = host.disk_threshold(12.5, (128, 128, 64))
    descs = [disk(thr, spikes=[(idxs[b], host.exp_f32(15.0))], wrap_alpha=0.5) for b in range(B_)]
    yg = chain(x, descs, True)
    assert rel_l2(y.cpu().numpy(), yg.cpu().numpy()) <= TOL
    ref2 = P.chain_127(x[2].cpu(), 12.5, idxs[2], 15.0, 0.5, 0.0, None)
    assert rel_l2(y[2].cpu().numpy(), ref2.numpy()) <= TOL


def test_spike_fast_path_matches_general_and_is_taken(cuda_device):
    """Spikes only (no mask, no wrap): two kernels (coefficient reduction + plane-wave axpy) instead of the FFT
    pipeline; same result as the general path and the oracle, min/max fused."""
    import ctypes as C
    from mvtb import _lib, functional as Fn, host
    from oracle import ref_port as P
    shape = (2, 240, 240, 155)
    x = P.synthetic_volume(21, shape)
    idx = [(120 + 40, 120 - 33, 77 + 20), (120 - 9, 120 + 2, 77 - 61)]
    descs = [host.make_desc(spikes=[(idx[c], host.exp_f32(15.0))]) for c in range(2)]
    xd = x.to(cuda_device)
    plan = Fn.get_plan(shape[1:], 2, cuda_device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    y, mm = Fn.kspace_chain(xd, 3, descs, want_minmax=True, vols_per_sample=2)
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    assert {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]} == {"k_spike_reduce", "k_spike_apply"}
    assert float(mm[0, 0]) == float(y.min()) and float(mm[0, 1]) == float(y.max())
    yg = chain(xd, descs, True)
    assert rel_l2(y.cpu().numpy(), yg.cpu().numpy()) <= TOL
    for c in range(2):
        ref = P.plane_wave_spike(x[c:c + 1], idx[c], 15.0)
        assert rel_l2(y[c:c + 1].cpu().numpy(), ref.numpy()) <= TOL


@pytest.mark.parametrize("shape,alpha", [((1, 240, 240, 155), 0.9), ((2, 128, 128, 64), 0.8), ((1, 240, 240, 155), 0.85)])
def test_bl_centred_mask_gibbs_noise_small_radius(cuda_device, shape, alpha):
    """GibbsNoise with alpha near 1 keeps a small centred ball: band-limited kernels, same numbers as the oracle."""
    import ctypes as C
    import filters_and_operators as F
    from mvtb import _lib, functional as Fn
    from oracle import ref_port as P
    x = P.synthetic_volume(12, shape)
    plan = Fn.get_plan(shape[1:], shape[0], cuda_device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    y = F.GibbsNoise(alpha)(x.to(cuda_device))
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    kinds = {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]}
    assert ("k_bl_fwd_h" in kinds or "k_bl_fwd_tc" in kinds) and "k_rows_fwd" not in kinds
    assert rel_l2(y.cpu().numpy(), P.gibbs_noise(x, alpha).numpy()) <= TOL
