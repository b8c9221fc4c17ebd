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
