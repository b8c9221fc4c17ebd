"""The tensor-core (tcgen05, 3xTF32) H-axis kernels of the band-limited path against the CUDA-core kernels and the
oracle: same Y layout, fp32-level accuracy (rel-L2 <= 1e-5 to the oracle, a few 1e-6 to the FFMA kernels), and no
bounded wait expired."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _run(x, descs, path, want_minmax=False):
    from mvtb import _lib, functional as Fn
    L = _lib.lib()
    plan = Fn.get_plan(x.shape[-3:], x.numel() // int(np.prod(x.shape[-3:])), x.device)
    _lib.check(L, L.mvtb_plan_set_path(plan, path))
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    try:
        out = Fn.kspace_chain(x, 3, descs, want_minmax=want_minmax)
        torch.cuda.synchronize()
        ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
        _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
        kinds = {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]}
        assert L.mvtb_plan_tc_status(plan) == 0, "a tensor-core kernel gave up on an mbarrier wait"
    finally:
        _lib.check(L, L.mvtb_plan_profile(plan, 0))
        _lib.check(L, L.mvtb_plan_set_path(plan, 0))
    return out, kinds


@pytest.mark.parametrize("shape,r", [((2, 1, 240, 240, 155), 12.5), ((3, 1, 128, 128, 64), 12.5), ((2, 1, 128, 128, 64), 25.0),
                                     ((5, 1, 64, 48, 40), 6.5), ((2, 2, 32, 36, 31), 3.0), ((1, 1, 240, 240, 155), 30.0),
                                     ((1, 1, 48, 50, 30), 9.0)])
def test_tc_forward_matches_cuda_cores_and_oracle(cuda_device, shape, r):
    from mvtb import _lib, host
    from oracle import ref_port as P
    xs = [P.synthetic_volume(50 + i, shape[1:]) for i in range(shape[0])]
    x = torch.stack(xs).to(cuda_device)
    d = [host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=host.disk_threshold(r, shape[-3:]))]
    y_tc, k_tc = _run(x, d, 5)
    y_cc, k_cc = _run(x, d, 4)
    assert "k_bl_fwd_tc" in k_tc and "k_bl_fwd_tc" not in k_cc and "k_bl_fwd_h" in k_cc
    assert rel_l2(y_tc.cpu().numpy(), y_cc.cpu().numpy()) <= 5e-6
    ref = P.fourier_disk_mask(xs[0], r, False).numpy()
    assert rel_l2(y_tc[0].cpu().numpy(), ref) <= TOL
    assert rel_l2(y_cc[0].cpu().numpy(), ref) <= TOL


def test_tc_chain127_with_spikes_and_minmax(cuda_device):
    from mvtb import _lib, functional as Fn, host
    from oracle import ref_port as P
    shape3 = (240, 240, 155)
    B_ = 5
    xs = [P.synthetic_volume(70 + i, (1,) + shape3) for i in range(B_)]
    x = torch.stack(xs).to(cuda_device)
    shell = host.ellipsoid_shell(shape3, 55., 55., 30.)
    thr = host.disk_threshold(12.5, shape3)
    descs, idxs = [], []
    for i in range(B_):
        idx = tuple(int(v) for v in shell[np.random.RandomState(i).randint(0, len(shell))]) if i != 2 else (123, 118, 80)
        idxs.append(idx)
        descs.append(host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=thr, spikes=[(idx, host.exp_f32(15.0 if i != 2 else 9.0))], wrap_alpha=0.5))
    (y, mm), kinds = _run(x, descs, 5, want_minmax=True)
    assert "k_bl_fwd_tc" in kinds
    for i in (0, 2):
        if i == 2:
            ref = P.chain_127(xs[i], 12.5, idxs[i], 9.0, 0.5, 0.0, None).numpy()         # spike inside the ball: well conditioned
        else:
            ref = P.chain_127_exact_phase(xs[i], 12.5, idxs[i], 15.0, 0.5).numpy()
        assert rel_l2(y[i].cpu().numpy(), ref) <= TOL
        assert float(mm[i, 0]) == float(y[i].min()) and float(mm[i, 1]) == float(y[i].max())


class _Path:
    """plan of (shape3, n) pinned to a path, with the per-kernel profile on; .kinds after the block"""

    def __init__(self, shape3, n, device, path):
        from mvtb import _lib, functional as Fn
        self.L, self.lib = _lib.lib(), _lib
        self.plan = Fn.get_plan(shape3, n, device)
        self.path = path
        self.kinds = set()

    def __enter__(self):
        self.lib.check(self.L, self.L.mvtb_plan_set_path(self.plan, self.path))
        self.lib.check(self.L, self.L.mvtb_plan_profile(self.plan, 1))
        return self

    def __exit__(self, *exc):
        torch.cuda.synchronize()
        ms, cn = (C.c_double * self.lib.K_KINDS)(), (C.c_int * self.lib.K_KINDS)()
        self.lib.check(self.L, self.L.mvtb_plan_profile_read(self.plan, ms, cn))
        self.kinds = {self.L.mvtb_kernel_name(k).decode() for k in range(self.lib.K_KINDS) if cn[k]}
        st = self.L.mvtb_plan_tc_status(self.plan)
        self.lib.check(self.L, self.L.mvtb_plan_profile(self.plan, 0))
        self.lib.check(self.L, self.L.mvtb_plan_set_path(self.plan, 0))
        assert st == 0, f"a tensor-core kernel gave up on an mbarrier wait (code {st})"
        return False


@pytest.mark.parametrize("shape,vps,p,r", [((6, 1, 64, 48, 40), 1, 0.05, 6.5), ((3, 2, 128, 128, 64), 2, 0.15, 12.5),
                                           ((5, 1, 32, 36, 31), 1, 0.35, 6.5), ((2, 4, 64, 64, 30), 4, 0.08, 6.5),
                                           ((3, 1, 48, 50, 30), 1, 0.05, 9.0), ((2, 1, 240, 240, 155), 1, 0.05, 12.5)])
def test_tc_inverse_with_select_in_the_stores(cuda_device, shape, vps, p, r):
    """Tensor-core path: compute-only (min, max) pass + (hit, coin) bits + store pass with the select applied ==
    the same chain followed by mvtb_salt_pepper_sparse_f32, bit for bit (same spans, same Philox counters), in place
    too; and the chain itself against the CUDA-core kernels."""
    from mvtb import _lib, functional as Fn, host
    from oracle import ref_port as P
    x = torch.stack([P.synthetic_volume(i, shape[1:]) for i in range(shape[0])]).to(cuda_device)
    thr = host.disk_threshold(r, shape[-3:])
    descs = []
    for b in range(shape[0]):
        idx = (shape[2] // 2 + 9 + b % 3, shape[3] // 2 - 11, shape[4] // 2 + 8)       # outside the ball: plane waves
        spikes = [(idx, host.exp_f32(9.0))]
        if b % 2:
            spikes.append(((shape[2] // 2 - 10, shape[3] // 2 + 9, shape[4] // 2 - 9), host.exp_f32(8.0)))
        d = host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=thr, spikes=spikes, wrap_alpha=0.25)
        descs.extend([d] * vps)
    n = shape[0] * vps
    with _Path(shape[-3:], n, cuda_device, 5) as t1:
        y3, mm = Fn.kspace_chain(x, 3, descs, want_minmax=True, vols_per_sample=vps)
        y3 = y3.clone()
    assert "k_bl_inv_tc" in t1.kinds and "k_bl_inv_h" not in t1.kinds and "k_bl_mm_tc" not in t1.kinds
    for s in range(shape[0]):
        assert float(mm[s, 0]) == float(y3[s].min()) and float(mm[s, 1]) == float(y3[s].max())
    want = Fn.salt_pepper(y3, p, seed=31, offset=12345, n_samples=shape[0], mm=mm, sparse=True)
    with _Path(shape[-3:], n, cuda_device, 5) as t2:
        got, mm2 = Fn.kspace_chain_sp(x, 3, descs, p, seed=31, offset=12345, vols_per_sample=vps)
        got = got.clone()
        z = x.clone()
        got2, _ = Fn.kspace_chain_sp(z, 3, descs, p, seed=31, offset=12345, vols_per_sample=vps, out=z)
    assert {"k_bl_mm_tc", "k_sp_bits", "k_bl_inv_tc"} <= t2.kinds and "k_bl_inv_sp" not in t2.kinds
    assert torch.equal(mm, mm2)
    assert torch.equal(got, want) and not torch.equal(got, y3)
    assert got2.data_ptr() == z.data_ptr() and torch.equal(z, want)
    with _Path(shape[-3:], n, cuda_device, 4) as t3:
        y_cc = Fn.kspace_chain(x, 3, descs, vols_per_sample=vps)
    assert "k_bl_inv_tc" not in t3.kinds
    assert rel_l2(y3.cpu().numpy(), y_cc.cpu().numpy()) <= 5e-6


@pytest.mark.parametrize("shape,vps,p,r", [((6, 1, 64, 48, 40), 1, 0.05, 6.5), ((3, 2, 128, 128, 64), 2, 0.15, 12.5),
                                           ((3, 1, 48, 50, 30), 1, 0.05, 9.0), ((5, 1, 240, 240, 155), 1, 0.05, 12.5)])
def test_tc_inverse_with_select_warps(cuda_device, monkeypatch, shape, vps, p, r):
    """MVTB_TC_INV=2: one persistent kernel stores the tensor-core inverse pass and, on four more warps, runs the select
    pass of every sample as soon as its tiles are counted complete == the two-call result, bit for bit."""
    from mvtb import _lib, functional as Fn, host
    from oracle import ref_port as P
    x = torch.stack([P.synthetic_volume(i, shape[1:]) for i in range(shape[0])]).to(cuda_device)
    thr = host.disk_threshold(r, shape[-3:])
    descs = []
    for b in range(shape[0]):
        idx = (shape[2] // 2 + 9 + b % 3, shape[3] // 2 - 11, shape[4] // 2 + 8)
        d = host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=thr, spikes=[(idx, host.exp_f32(9.0))], wrap_alpha=0.25)
        descs.extend([d] * vps)
    n = shape[0] * vps
    with _Path(shape[-3:], n, cuda_device, 5):
        y3, mm = Fn.kspace_chain(x, 3, descs, want_minmax=True, vols_per_sample=vps)
        y3 = y3.clone()
    want = Fn.salt_pepper(y3, p, seed=77, offset=5, n_samples=shape[0], mm=mm, sparse=True)
    monkeypatch.setenv("MVTB_TC_INV", "2")
    Fn._destroy_plans()
    try:
        L = _lib.lib()
        plan = Fn.get_plan(shape[-3:], n, cuda_device)
        _lib.check(L, L.mvtb_plan_profile(plan, 1))
        got, mm2 = Fn.kspace_chain_sp(x, 3, descs, p, seed=77, offset=5, vols_per_sample=vps)
        torch.cuda.synchronize()
        ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
        _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
        kinds = {L.mvtb_kernel_name(k).decode() for k in range(_lib.K_KINDS) if cn[k]}
        assert L.mvtb_plan_tc_status(plan) == 0
        got = got.clone()
    finally:
        Fn._destroy_plans()
    assert "k_bl_inv_tc" in kinds and not ({"k_bl_mm_tc", "k_sp_bits", "k_bl_inv_sp", "k_bl_inv_h"} & kinds)
    assert torch.equal(mm, mm2)
    assert torch.equal(got, want) and not torch.equal(got, y3)


@pytest.mark.parametrize("shape,r,expect_tc", [((70, 1, 16, 20, 12), 3.0, True),       # more volumes than a chunk, the smallest H
                                               ((3, 1, 32, 15, 7), 3.0, False),        # W*D % 4 != 0: no TMA rows, FFMA kernel
                                               ((3, 1, 40, 36, 32), 6.5, False)])      # H % 16 != 0
def test_default_path_picks_the_forward_kernel_by_shape(cuda_device, shape, r, expect_tc):
    """The automatic path runs the tensor-core forward kernel exactly when its layout conditions hold, and either way
    matches the oracle."""
    from mvtb import _lib, functional as Fn, host
    from oracle import ref_port as P
    xs = [P.synthetic_volume(300 + i, shape[1:]) for i in range(min(shape[0], 2))]
    x = torch.stack([xs[i % len(xs)] for i in range(shape[0])]).to(cuda_device)
    d = [host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=host.disk_threshold(r, shape[-3:]))]
    with _Path(shape[-3:], shape[0], cuda_device, 0) as t:
        y = Fn.kspace_chain(x, 3, d).clone()
    assert ("k_bl_fwd_tc" in t.kinds) == expect_tc and ("k_bl_fwd_h" in t.kinds) == (not expect_tc)
    ref = P.fourier_disk_mask(xs[0], r, False).numpy()
    assert rel_l2(y[0].cpu().numpy(), ref) <= TOL
    assert torch.equal(y[len(xs)], y[0])                     # same input volume, another chunk slot / tile range: same bits
