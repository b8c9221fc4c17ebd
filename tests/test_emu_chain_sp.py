"""mvtb_kspace_chain_sp_f32 (chain + salt-and-pepper in one call; the select pass runs inside the persistent
inverse kernel on the band-limited path) through the DEBUG emulator: bit-identical to the two separate calls,
for every queue shape the host can build.  The CUDA evidence is tests/test_gpu_chain_sp.py."""
import ctypes as C
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")

from cuemu import emu  # noqa: E402
from mvtb import _lib as B, host  # noqa: E402
from oracle import ref_port as P  # noqa: E402


def two_calls(x, descs, vps, p, seed, offset, chunk=4, general=False):
    L = emu.lib()
    plan = emu.Plan(x.shape[-3:], chunk)
    if general:
        B.check(L, L.mvtb_plan_set_path(plan.h, 1))
    nvol = int(np.prod(x.shape[:-3]))
    y = np.empty_like(x)
    mm = np.zeros(2 * (nvol // vps), dtype=np.float32)
    B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(y), nvol, host.desc_array(descs), len(descs),
                                       emu.ptr(mm), vps, None))
    y3 = y.copy()
    tab = np.zeros(B.SP_BLOCK, dtype=np.uint32)
    B.check(L, L.mvtb_salt_pepper_sparse_f32(emu.ptr(y), y.size // (nvol // vps), nvol // vps, seed, offset, C.c_float(p),
                                             emu.ptr(mm), emu.ptr(tab), None))
    return y3, y, mm


def one_call(x, descs, vps, p, seed, offset, chunk=4, general=False):
    L = emu.lib()
    plan = emu.Plan(x.shape[-3:], chunk)
    if general:
        B.check(L, L.mvtb_plan_set_path(plan.h, 1))
    nvol = int(np.prod(x.shape[:-3]))
    y = np.empty_like(x)
    mm = np.zeros(2 * (nvol // vps), dtype=np.float32)
    B.check(L, L.mvtb_plan_profile(plan.h, 1))
    B.check(L, L.mvtb_kspace_chain_sp_f32(plan.h, emu.ptr(x), emu.ptr(y), nvol, host.desc_array(descs), len(descs),
                                          emu.ptr(mm), vps, C.c_float(p), seed, offset, None))
    ms, cn = (C.c_double * B.K_KINDS)(), (C.c_int * B.K_KINDS)()
    B.check(L, L.mvtb_plan_profile_read(plan.h, ms, cn))
    kinds = {L.mvtb_kernel_name(k).decode() for k in range(B.K_KINDS) if cn[k]}
    return y, mm, kinds


def descs_for(shape, r, n, spike=True, alpha=0.5):
    thr = host.disk_threshold(r, shape)
    out = []
    for i in range(n):
        sp = [((shape[0] // 2 + 6 + i % 2, shape[1] // 2 - 4, shape[2] // 2 + 3), 3.0 + i)] if spike else []
        out.append(host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=thr, spikes=sp, wrap_alpha=alpha))
    return out


@pytest.mark.parametrize("hs,lag,spread", [(None, None, None), (1, 0, 100), (2, 3, 50), (4, 40, 100), (3, 100, 10)])
@pytest.mark.parametrize("shape,vps,p", [((6, 16, 12, 10), 1, 0.05), ((6, 16, 12, 10), 2, 0.3), ((4, 24, 9, 14), 4, 0.12)])
def test_fused_is_bit_identical_to_two_calls(monkeypatch, shape, vps, p, hs, lag, spread):
    if hs is not None:
        monkeypatch.setenv("MVTB_IS_HS", str(hs))
        monkeypatch.setenv("MVTB_IS_LAG", str(lag))
        monkeypatch.setenv("MVTB_IS_SPREAD", str(spread))
    x = P.synthetic_volume(3, shape).numpy()
    d = descs_for(shape[-3:], 3.5, shape[0])
    for k in range(0, shape[0], vps):                       # one spike location per sample, as the 127 chain has
        for c in range(vps):
            d[k + c] = d[k]
    y3, want, mm_want = two_calls(x, d, vps, p, 99, 7)
    got, mm, kinds = one_call(x, d, vps, p, 99, 7)
    assert "k_bl_inv_sp" in kinds and "k_bl_inv_h" not in kinds
    assert np.array_equal(mm, mm_want)
    assert np.array_equal(got, want)
    assert not np.array_equal(got, y3)                      # the select pass did something


def test_chunks_of_whole_samples(monkeypatch):
    """More samples than one launch holds: Philox counters and minmax slots continue across launches."""
    monkeypatch.setenv("MVTB_IS_CHUNK", "4")
    shape = (10, 16, 12, 10)
    x = P.synthetic_volume(4, shape).numpy()
    d = descs_for(shape[-3:], 3.5, 10)
    for vps in (1, 2):
        dd = list(d)
        for k in range(0, 10, vps):
            for c in range(vps):
                dd[k + c] = dd[k]
        _, want, mm_want = two_calls(x, dd, vps, 0.2, 5, 1000)
        got, mm, kinds = one_call(x, dd, vps, 0.2, 5, 1000)
        assert "k_bl_inv_sp" in kinds
        assert np.array_equal(got, want) and np.array_equal(mm, mm_want)


def test_in_place_and_shared_descriptor():
    shape = (3, 16, 12, 10)
    x = P.synthetic_volume(6, shape).numpy()
    d = descs_for(shape[-3:], 3.5, 1, spike=False, alpha=None)
    _, want, _ = two_calls(x, d, 1, 0.1, 1, 0)
    L = emu.lib()
    plan = emu.Plan(shape[-3:], 4)
    y = x.copy()
    mm = np.zeros(6, dtype=np.float32)
    B.check(L, L.mvtb_kspace_chain_sp_f32(plan.h, emu.ptr(y), emu.ptr(y), 3, host.desc_array(d), 1, emu.ptr(mm), 1,
                                          C.c_float(0.1), 1, 0, None))
    assert np.array_equal(y, want)


@pytest.mark.parametrize("case", ["general", "odd_h", "p0", "spike_only"])
def test_unfused_paths_give_the_same_result(case):
    """Shapes / paths the persistent kernel does not cover run the select pass as its own kernel: same output."""
    shape = (2, 16, 12, 10) if case != "odd_h" else (2, 15, 12, 10)
    x = P.synthetic_volume(8, shape).numpy()
    if case == "spike_only":
        d = [host.make_desc(spikes=[((3, 4, 5), 4.0)])]
    else:
        d = descs_for(shape[-3:], 3.5, 1, spike=False)
    p = 0.0 if case == "p0" else 0.25
    y3, want, _ = two_calls(x, d, 1, p, 11, 3, general=(case == "general"))
    got, _, kinds = one_call(x, d, 1, p, 11, 3, general=(case == "general"))
    assert "k_bl_inv_sp" not in kinds
    assert np.array_equal(got, want)
    if p == 0.0:
        assert np.array_equal(got, y3)


def test_argument_errors():
    L = emu.lib()
    shape = (3, 16, 12, 10)
    x = P.synthetic_volume(6, shape).numpy()
    d = descs_for(shape[-3:], 3.5, 1, spike=False)
    plan = emu.Plan(shape[-3:], 4)
    y = np.empty_like(x)
    mm = np.zeros(6, dtype=np.float32)
    arr = host.desc_array(d)
    assert L.mvtb_kspace_chain_sp_f32(plan.h, emu.ptr(x), emu.ptr(y), 3, arr, 1, None, 1, C.c_float(0.1), 1, 0, None) == B.MVTB_EINVAL
    assert L.mvtb_kspace_chain_sp_f32(plan.h, emu.ptr(x), emu.ptr(y), 3, arr, 1, emu.ptr(mm), 2, C.c_float(0.1), 1, 0, None) == B.MVTB_EINVAL
    assert L.mvtb_kspace_chain_sp_f32(plan.h, emu.ptr(x), emu.ptr(y), 3, arr, 1, emu.ptr(mm), 1, C.c_float(1.5), 1, 0, None) == B.MVTB_EINVAL


def test_spike_on_the_boundary_shell_of_a_centred_mask():
    """GibbsNoise then KSpaceSpikeNoise as ONE descriptor, with the spike on a bin where the centred mask keeps f_s
    but not -f_s (even axes): the spike stage must read M_eff(f_s) K = K / 2, what the mask stage's real output
    holds there.  Both chain paths against the oracle's two sequential calls."""
    import torch
    from conftest import rel_l2
    shape = (1, 8, 8, 6)
    x = P.synthetic_volume(2, shape)
    x = x + 0.3 * torch.randn(shape, generator=torch.Generator().manual_seed(1))   # energy in every bin
    alpha = 0.45
    thr = host.gibbs_threshold(alpha, shape[1:])
    mask = P.gibbs_mask(shape[1:], alpha)
    cand = [(i, j, k) for i in range(8) for j in range(8) for k in range(6)
            if mask[i, j, k] and not mask[(8 - i) % 8, (8 - j) % 8, (6 - k) % 6]]
    assert cand, "no asymmetric shell bin for this alpha"
    idx = cand[len(cand) // 2]
    ref = P.kspace_spike(P.gibbs_noise(x, alpha), idx, 6.0).numpy()
    d = host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=3, mask_thresh=thr, spikes=[(idx, host.exp_f32(6.0))])
    for general in (True, False):
        L = emu.lib()
        plan = emu.Plan(shape[1:], 2)
        B.check(L, L.mvtb_plan_set_path(plan.h, 1 if general else 0))
        xn = x.numpy()
        y = np.empty_like(xn)
        B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(xn), emu.ptr(y), 1, host.desc_array([d]), 1, None, 1, None))
        assert rel_l2(y, ref) <= 1e-5, (general, rel_l2(y, ref))
