"""Band-limited (pruned-DFT) path of the chain, run through the DEBUG emulator, against the
general FFT path and the oracle.  Debug scaffolding for a GPU-less container, like
tests/test_emu_kernels.py; the parity evidence for the CUDA build is tests/test_gpu_bandlimited.py."""
import shutil

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")

from cuemu import emu  # noqa: E402
from mvtb import _lib as B, host  # noqa: E402
from oracle import ref_port as P  # noqa: E402

TOL = 1e-5


@pytest.fixture(autouse=True, params=[True, False], ids=["fusedmid", "splitmid"])
def fused(request, monkeypatch):
    """Every test runs with the W/D stage as one kernel (default) and as three (MVTB_NO_FUSEMID, read at plan create)."""
    if not request.param:
        monkeypatch.setenv("MVTB_NO_FUSEMID", "1")
    return request.param


def run(x, descs, general, chunk=4, vps=None):
    L = emu.lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    plan = emu.Plan(x.shape[-3:], chunk)
    B.check(L, L.mvtb_plan_set_path(plan.h, 1 if general else 0))
    n0 = L.mvtb_launch_count()
    y = np.empty_like(x)
    nvol = int(np.prod(x.shape[:-3]))
    mm = np.zeros(2 * ((nvol + vps - 1) // vps), dtype=np.float32) if vps else None
    B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(y), nvol, host.desc_array(descs), len(descs),
                                       emu.ptr(mm), vps or 1, None))
    return y, int(L.mvtb_launch_count() - n0), mm


def disk(thr, **kw):
    return host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=thr, **kw)


@pytest.mark.parametrize("shape,r", [((2, 16, 12, 8), 2.5), ((1, 32, 32, 16), 4.0), ((1, 9, 15, 25), 2.5),
                                     ((3, 12, 10, 31), 3.0), ((1, 30, 26, 27), 12.5), ((1, 44, 40, 39), 18.5),
                                     ((1, 64, 51, 52), 25.0), ((1, 63, 64, 65), 30.5)])
def test_bl_matches_oracle_and_general(shape, r):
    x = P.synthetic_volume(1, shape).numpy()
    d = disk(host.disk_threshold(r, shape[-3:]))
    yb, nb, _ = run(x, [d], False)
    yg, ng, _ = run(x, [d], True)
    ref = P.fourier_disk_mask(torch.from_numpy(x), r).numpy()
    assert rel_l2(yb, ref) <= TOL and rel_l2(yg, ref) <= TOL and rel_l2(yb, yg) <= TOL


def test_bl_is_actually_taken(fused):
    """3 band-limited launches (5 with the W/D stage split by MVTB_NO_FUSEMID), none of the general kernels' kinds,
    and both variants agree with the oracle."""
    xt = P.synthetic_volume(1, (2, 16, 12, 8))
    x = xt.numpy()
    L = emu.lib()
    plan = emu.Plan((16, 12, 8), 4)
    B.check(L, L.mvtb_plan_profile(plan.h, 1))
    y = np.empty_like(x)
    d = disk(host.disk_threshold(2.5, (16, 12, 8)))
    B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(y), 2, host.desc_array([d]), 1, None, 1, None))
    import ctypes as C
    ms, cn = (C.c_double * B.K_KINDS)(), (C.c_int * B.K_KINDS)()
    B.check(L, L.mvtb_plan_profile_read(plan.h, ms, cn))
    kinds = {L.mvtb_kernel_name(k).decode(): cn[k] for k in range(B.K_KINDS) if cn[k]}
    want = {"k_bl_fwd_h": 1, "k_bl_mid": 1, "k_bl_inv_h": 1}
    if not fused:
        want.update({"k_bl_fwd_w": 1, "k_bl_inv_w": 1})
    assert kinds == want
    ref = torch.stack([P.fourier_disk_mask(xt[c], 2.5) for c in range(2)]).numpy()
    assert rel_l2(y, ref) <= TOL


def test_bl_spikes_wrap_per_volume_descs_and_minmax():
    shape = (3, 20, 18, 15)
    x = P.synthetic_volume(2, shape).numpy()
    thr = host.disk_threshold(3.2, shape[-3:])
    amp = host.exp_f32(6.0)
    descs = [disk(thr, spikes=[((11, 10, 8), amp)], wrap_alpha=0.5),                               # inside the ball
             disk(thr, spikes=[((2, 15, 1), amp), ((10, 9, 7), 0.5 * amp)], wrap_alpha=0.25),      # plane wave + DC bin
             disk(thr, spikes=[((0, 0, 0), amp), ((13, 9, 7), amp)])]                              # Nyquist corner + in box, outside ball
    yb, nb, mb = run(x, descs, False, vps=1)
    yg, ng, mg = run(x, descs, True, vps=1)
    for i in range(3):
        assert rel_l2(yb[i], yg[i]) <= TOL
        assert mb[2 * i] == yb[i].min() and mb[2 * i + 1] == yb[i].max()
    assert np.allclose(mb, mg, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("chain127_"))
def test_bl_chain127_golden(name):
    m, z = load_golden(name)
    x = z["x"]
    thr = host.disk_threshold(m["r"], x.shape[-3:])
    amp = host.exp_f32(m["intensity"])
    d = disk(thr, spikes=[(m["idx"], amp)], wrap_alpha=m["alpha"])
    yb, nb, _ = run(x, [d], False)
    yg, ng, _ = run(x, [d], True)
    assert rel_l2(yb, yg) <= TOL                   # identical treatment of the zeroed-bin phase in both paths
    if m["spike_in_ball"]:
        assert rel_l2(yb, z["y3"]) <= TOL


def test_bl_quad_symmetry_kernels_match_pair_kernels():
    """H % 4 == 0 uses the four-rows-per-table-row kernels; MVTB_PATH_BL_PAIRS (2) forces the pair kernels."""
    shape = (2, 24, 14, 11)
    x = P.synthetic_volume(4, shape).numpy()
    thr = host.disk_threshold(3.3, shape[-3:])
    amp = host.exp_f32(5.0)
    descs = [disk(thr, spikes=[((3, 9, 2), amp), ((20, 1, 10), amp)], wrap_alpha=0.5),     # odd and even f_h plane waves
             disk(thr, spikes=[((12, 7, 5), amp)])]
    L = emu.lib()
    outs = []
    for path in (0, 2, 1):
        plan = emu.Plan(shape[-3:], 4)
        B.check(L, L.mvtb_plan_set_path(plan.h, path))
        y = np.empty_like(x)
        mm = np.zeros(4, dtype=np.float32)
        B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(y), 2, host.desc_array(descs), 2, emu.ptr(mm), 1, None))
        outs.append((y, mm))
    assert rel_l2(outs[0][0], outs[1][0]) <= TOL and rel_l2(outs[0][0], outs[2][0]) <= TOL
    assert np.allclose(outs[0][1], outs[2][1], rtol=1e-5, atol=1e-6)


def test_bl_many_volumes_chunked_workspace():
    shape = (11, 14, 12, 9)
    x = P.synthetic_volume(3, shape).numpy()
    thr = host.disk_threshold(2.2, shape[-3:])
    descs = [disk(thr, spikes=[((i % 14, (3 * i) % 12, (5 * i) % 9), 50.0 + i)]) for i in range(11)]
    yb, _, mb = run(x, descs, False, chunk=1, vps=4)
    yg, _, mg = run(x, descs, True, chunk=1, vps=4)
    assert rel_l2(yb, yg) <= TOL
    assert np.allclose(mb, mg, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape,alpha", [((2, 16, 12, 8), 0.8), ((1, 15, 12, 9), 0.7), ((1, 9, 15, 7), 0.75), ((3, 20, 18, 15), 0.9)])
def test_bl_centred_masks_gibbs_noise(shape, alpha, fused):
    """GibbsNoise / GibbsNoiseLayer masks (centre (N-1)/2, M_eff in {0, 1/2, 1} on even axes) with a small radius
    take the band-limited path too: same numbers as the general path and the oracle."""
    xt = P.synthetic_volume(7, shape)
    x = xt.numpy()
    d = host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=3, mask_thresh=host.gibbs_threshold(alpha, shape[1:]))
    yb, nb, _ = run(x, [d], general=False)
    yg, ng, _ = run(x, [d], general=True)
    ref = P.gibbs_noise(xt, alpha).numpy()
    assert nb == (3 if fused else 5) and ng == 5      # 3 launches: only the band-limited path has that few
    assert rel_l2(yb, ref) <= TOL and rel_l2(yg, ref) <= TOL


@pytest.mark.parametrize("spike_d", [None, 38])
def test_general_path_skips_tiles_the_mask_removes(spike_d):
    """General FFT path with a small ball on a long last axis (D = 40: half-spectrum bins 16..20 form a tile that the
    mask removes whatever the other frequencies are): the skipped tiles must come out as exact zeros in k-space,
    i.e. the result still equals the band-limited path and the oracle -- also when a spike sits in such a tile."""
    shape = (3, 8, 6, 40)
    xt = P.synthetic_volume(13, shape)
    x = xt.numpy()
    thr = host.disk_threshold(3.2, shape[1:])
    sp = [((4 + 1, 3, spike_d), host.exp_f32(5.0))] if spike_d is not None else []
    d = disk(thr, spikes=sp)
    yg, _, _ = run(x, [d], general=True)
    yb, _, _ = run(x, [d], general=False)
    assert rel_l2(yg, yb) <= TOL
    if spike_d is None:
        ref = torch.stack([P.fourier_disk_mask(xt[c], 3.2) for c in range(3)]).numpy()
        assert rel_l2(yg, ref) <= TOL
