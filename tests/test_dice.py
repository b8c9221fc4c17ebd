"""Dice loss / metric: restatement known answers, emulated kernel sums against numpy."""
import shutil

import numpy as np
import pytest
import torch

from oracle import monai_losses as ML


def test_restatement_known_answers():
    t = torch.zeros(1, 1, 2, 2, 2)
    t[0, 0, 0] = 1.0
    big = torch.where(t > 0, torch.tensor(30.0), torch.tensor(-30.0))          # sigmoid -> exactly 1 / ~0
    assert float(ML.dice_loss(big, t)) < 1e-6                                   # perfect prediction
    assert abs(float(ML.dice_loss(-big, t)) - 1.0) < 1e-5                       # disjoint prediction
    v, nn_ = ML.dice_metric(ML.post_trans(big), t)
    assert float(v) == 1.0 and float(nn_) == 1.0
    v, nn_ = ML.dice_metric(ML.post_trans(big), torch.zeros_like(t))            # empty ground truth: NaN -> not counted
    assert float(v) == 0.0 and float(nn_) == 0.0
    half = torch.zeros(1, 1, 4)
    tt = torch.tensor([[[1.0, 1.0, 0.0, 0.0]]])
    # p = 0.5 everywhere: I = 1, sum p^2 = 1, sum t^2 = 2 -> 1 - 2/3
    assert abs(float(ML.dice_loss(half, tt)) - (1 - (2 + 1e-5) / (3 + 1e-5))) < 1e-6


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")
@pytest.mark.parametrize("from_logits", [1, 0])
def test_emulated_sums_and_grad(from_logits):
    from cuemu import emu
    from mvtb import _lib as B
    L = emu.lib()
    rng = np.random.RandomState(3)
    nv, n = 3, 7 * 9 * 5
    x = rng.randn(nv, n).astype(np.float32) * 2
    if not from_logits:
        x = (1 / (1 + np.exp(-x))).astype(np.float32)
    t = (rng.rand(nv, n) < 0.3).astype(np.float32)
    sums = np.zeros((nv, 6))
    scratch = np.zeros(int(L.mvtb_dice_scratch_bytes(nv)), dtype=np.uint8)
    B.check(L, L.mvtb_dice_sums_f32(emu.ptr(x), emu.ptr(t), n, nv, from_logits, emu.ptr(sums), emu.ptr(scratch), None))
    p = (1 / (1 + np.exp(-x.astype(np.float64)))) if from_logits else x.astype(np.float64)
    q = (p >= 0.5).astype(np.float64)
    want = np.stack([(t * p).sum(1), (p * p).sum(1), (t * t).sum(1), (t * q).sum(1), q.sum(1), t.sum(1)], axis=1)
    assert np.allclose(sums, want, rtol=1e-6)
    coef = rng.randn(nv, 2).astype(np.float32)
    g = np.zeros_like(x)
    B.check(L, L.mvtb_dice_grad_f32(emu.ptr(x), emu.ptr(t), n, nv, from_logits, emu.ptr(coef), emu.ptr(g), None))
    d = coef[:, :1] * t + coef[:, 1:] * p
    wantg = d * (p * (1 - p)) if from_logits else d
    assert np.allclose(g, wantg, rtol=2e-5, atol=1e-6)
