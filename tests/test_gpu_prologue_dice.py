"""SURVEY 8(f) ranks 1 and 2 on the GPU: intensity prologue (alone, and applied on load by the chain) and the fused
Dice loss / metric, against the MONAI 0.5 restatements under oracle/."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _brainish(seed, shape):
    from oracle import ref_port as P
    x = P.synthetic_volume(seed, shape)
    return (x * 120.0 + (x != 0) * 400.0).to(torch.float32)       # raw-MRI-like: positive inside the support, 0 outside


@pytest.mark.parametrize("shape", [(1, 240, 240, 155), (4, 128, 128, 64), (2, 33, 17, 29)])
def test_prologue_matches_monai_restatement(cuda_device, shape):
    from mvtb import functional as Fn
    from oracle import monai_intensity as M
    x = _brainish(7, shape)
    ref = M.prologue(x.numpy(), 0.07, -0.03)
    xd = x.to(cuda_device)
    abt, stats = Fn.intensity_coeffs(xd, shape[0], scale=1.07, shift=-0.03, want_stats=True)
    y = Fn.intensity_affine(xd, abt)
    assert rel_l2(y.cpu().numpy(), ref) <= 1e-5
    for c in range(shape[0]):                             # masked mean / std in float64
        m = x[c] != 0
        v = x[c][m].double()
        assert float(stats[c, 0]) == float(m.sum())
        assert abs(float(stats[c, 1]) - float(v.mean())) <= 1e-9 * abs(float(v.mean()))
        assert abs(float(stats[c, 2]) - float(v.std(unbiased=False))) <= 1e-9 * float(v.std(unbiased=False))
    assert torch.equal(y.cpu()[x == 0], torch.full_like(y.cpu()[x == 0], np.float32(-0.03)))
    # drop-in classes, same draws as the restatement
    from mvtb import intensity as I
    tr = I.IntensityPrologued("image", factors=0.1, offsets=0.1, prob=0.5)
    tr.scale.set_random_state(seed=11)
    tr.shift.set_random_state(seed=12)
    out = tr({"image": x})["image"]
    f, o = M.draw_scale_shift(np.random.RandomState(11), np.random.RandomState(12))
    assert rel_l2(out.numpy(), M.prologue(x.numpy(), f, o)) <= 1e-5
    n1 = I.NormalizeIntensityd("image", nonzero=True, channel_wise=True)({"image": x})["image"]
    assert rel_l2(n1.numpy(), M.normalize_intensity(x.numpy())) <= 1e-5


def test_chain_applies_prologue_on_load(cuda_device):
    """prologue -> chain-127 -> sparse S&P as one library call == the three steps one after the other."""
    import ctypes as C
    from mvtb import _lib, functional as Fn, host
    shape3 = (240, 240, 155)
    B_ = 6
    x = torch.stack([_brainish(20 + b, (1,) + shape3) for b in range(B_)]).to(cuda_device)
    thr = host.disk_threshold(12.5, shape3)
    shell = host.ellipsoid_shell(shape3, 55., 55., 30.)
    descs = [host.make_desc(mask_kind=_lib.MASK_DISK, mask_ndim=3, mask_thresh=thr, wrap_alpha=0.5,
                            spikes=[(tuple(int(v) for v in shell[np.random.RandomState(b).randint(0, len(shell))]), host.exp_f32(15.0))])
             for b in range(B_)]
    sc = torch.linspace(0.92, 1.08, B_)
    sh = torch.linspace(-0.1, 0.1, B_)
    abt = Fn.intensity_coeffs(x, B_, scale=sc, shift=sh)
    xa = Fn.intensity_affine(x, abt)
    want3 = Fn.kspace_chain(xa, 3, descs)
    plan = Fn.get_plan(shape3, B_, cuda_device)
    L = _lib.lib()
    _lib.check(L, L.mvtb_plan_profile(plan, 1))
    got3 = Fn.kspace_chain_ex(x, 3, descs, pre_abt=abt)
    torch.cuda.synchronize()
    ms, cn = (C.c_double * _lib.K_KINDS)(), (C.c_int * _lib.K_KINDS)()
    _lib.check(L, L.mvtb_plan_profile_read(plan, ms, cn))
    _lib.check(L, L.mvtb_plan_profile(plan, 0))
    assert cn[5] + cn[14] == 1                            # k_bl_fwd_h / k_bl_fwd_tc: the map rode on the forward kernel, no extra pass
    assert rel_l2(got3.cpu().numpy(), want3.cpu().numpy()) <= 2e-6
    want4, _ = Fn.kspace_chain_sp(xa, 3, descs, 0.05, seed=3, offset=0)
    got4, mm = Fn.kspace_chain_ex(x, 3, descs, pre_abt=abt, sp=(0.05, 3, 0))
    hit_w, hit_g = want4 != want3, got4 != got3
    assert torch.equal(hit_w, hit_g)                      # same coordinates; values are min/2, max/2 of (slightly different) outputs
    # and against the restatement end to end for one volume (stage-wise tolerances apply: compare the k-space part)
    from oracle import monai_intensity as M, ref_port as P
    ref_in = torch.from_numpy(M.prologue(x[1].cpu().numpy(), float(sc[1]) - 1.0, float(sh[1])))
    idx = tuple(int(v) for v in shell[np.random.RandomState(1).randint(0, len(shell))])
    ref = P.chain_127_exact_phase(ref_in, 12.5, idx, 15.0, 0.5).numpy()
    assert rel_l2(got3[1].cpu().numpy(), ref) <= 1e-5


@pytest.mark.parametrize("shape", [(2, 1, 128, 128, 64), (3, 2, 40, 33, 21)])
def test_dice_loss_and_metric_match_monai_restatement(cuda_device, shape):
    from mvtb import losses as LS
    from oracle import monai_losses as ML
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(shape, generator=g) * 3).to(cuda_device)
    target = (torch.rand(shape, generator=g) < 0.2).float().to(cuda_device)
    target[0, 0] = 0                                      # an empty ground truth: NaN in the metric, excluded from the mean
    lg = logits.clone().requires_grad_(True)
    loss = LS.DiceLoss(to_onehot_y=False, sigmoid=True, squared_pred=True)(lg, target)
    loss.backward()
    lr = logits.clone().requires_grad_(True)
    ref = ML.dice_loss(lr, target)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-6
    assert rel_l2(lg.grad.cpu().numpy(), lr.grad.cpu().numpy()) <= 1e-5
    m = LS.DiceMetric(include_background=True, reduction="mean")
    v, nn_ = m(ML.post_trans(logits), target)
    rv, rn = ML.dice_metric(ML.post_trans(logits), target)
    assert abs(float(v) - float(rv)) <= 1e-6 and float(nn_) == float(rn)
    v2, nn2 = m.from_logits(logits, target)
    assert abs(float(v2) - float(rv)) <= 1e-6 and float(nn2) == float(rn)
    l3, (v3, _) = LS.dice_loss_and_metric(logits, target)
    assert abs(float(l3) - float(ref)) <= 1e-6 and abs(float(v3) - float(rv)) <= 1e-6
