"""Dice loss and Dice metric of the reference's training / evaluation loops as fused CUDA reductions
(csrc/dice.cu): drop-ins for monai.losses.DiceLoss(to_onehot_y=False, sigmoid=True, squared_pred=True) and
monai.metrics.DiceMetric(include_background=True, reduction="mean") as the scripts construct them
(10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:216, 266-283).  One pass over (logits, target)
yields the six per-volume sums both need; the loss has an analytic backward (one more pass).  No CPU fallback."""
import ctypes as C
from typing import Tuple

import torch
import torch.nn as nn

from . import _lib
from .functional import _ptr, _stream, require_cuda


def dice_sums(x: torch.Tensor, target: torch.Tensor, from_logits: bool = True) -> torch.Tensor:
    """(B, C, 6) float64: sum t p, sum p^2, sum t^2, sum t q, sum q, sum t over the spatial axes (p = sigmoid(x) or x,
    q = [p >= 0.5])."""
    require_cuda()
    if x.shape != target.shape or x.dim() < 3:
        raise AssertionError(f"ground truth has differing shape ({tuple(target.shape)}) from input ({tuple(x.shape)})")
    if not x.is_cuda or not target.is_cuda:
        raise ValueError("mvtb.losses run on CUDA tensors only")
    x = x.detach().to(torch.float32).contiguous()
    t = target.detach().to(torch.float32).contiguous()
    L = _lib.lib()
    nv = int(x.shape[0] * x.shape[1])
    sums = torch.empty((x.shape[0], x.shape[1], 6), dtype=torch.float64, device=x.device)
    if nv == 0:
        return sums
    scratch = torch.empty(max(int(L.mvtb_dice_scratch_bytes(nv)), 8), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.mvtb_dice_sums_f32(_ptr(x), _ptr(t), x.numel() // nv, nv, 1 if from_logits else 0, _ptr(sums), _ptr(scratch), _stream(x.device))
    _lib.check(L, rc)
    return sums


class _DiceLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, sigmoid, squared_pred, smooth_nr, smooth_dr, reduction):
        if not squared_pred:
            raise NotImplementedError("DiceLoss: only squared_pred=True (the reference's setting) is built")
        s = dice_sums(x, target, from_logits=sigmoid)
        inter, den = s[..., 0], s[..., 1] + s[..., 2]
        f = 1.0 - (2.0 * inter + smooth_nr) / (den + smooth_dr)
        ctx.save_for_backward(x, target, inter, den)
        ctx.cfg = (sigmoid, smooth_nr, smooth_dr, reduction)
        if reduction == "mean":
            out = f.mean()
        elif reduction == "sum":
            out = f.sum()
        else:
            out = f
        return out.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        x, target, inter, den = ctx.saved_tensors
        sigmoid, smooth_nr, smooth_dr, reduction = ctx.cfg
        nv = inter.numel()
        if reduction == "mean":
            gv = (g.to(torch.float64) / nv).expand_as(inter)
        elif reduction == "sum":
            gv = g.to(torch.float64).expand_as(inter)
        else:
            gv = g.to(torch.float64)
        # f = 1 - (2 I + e) / (D + e'),  I = sum t p,  D = sum t^2 + sum p^2:  df/dp_i = -2 t_i / (D + e') + 2 p_i (2 I + e) / (D + e')^2
        dd = den + smooth_dr
        coef = torch.stack([(-2.0 / dd) * gv, (2.0 * (2.0 * inter + smooth_nr) / (dd * dd)) * gv], dim=-1).to(torch.float32).contiguous()
        xf = x.detach().to(torch.float32).contiguous()
        tf = target.detach().to(torch.float32).contiguous()
        grad = torch.empty_like(xf)
        L = _lib.lib()
        with torch.cuda.device(xf.device):
            rc = L.mvtb_dice_grad_f32(_ptr(xf), _ptr(tf), xf.numel() // max(nv, 1), nv, 1 if sigmoid else 0, _ptr(coef), _ptr(grad), _stream(xf.device))
        _lib.check(L, rc)
        return grad.to(x.dtype), None, None, None, None, None, None


class DiceLoss(nn.Module):
    """monai.losses.DiceLoss for the reference's configuration; other options raise NotImplementedError."""

    def __init__(self, include_background: bool = True, to_onehot_y: bool = False, sigmoid: bool = False, softmax: bool = False,
                 other_act=None, squared_pred: bool = False, jaccard: bool = False, reduction: str = "mean",
                 smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False) -> None:
        super().__init__()
        if softmax or other_act is not None or to_onehot_y or jaccard or batch or not include_background:
            raise NotImplementedError("mvtb DiceLoss covers DiceLoss(to_onehot_y=False, sigmoid=..., squared_pred=True) as the reference uses it")
        if reduction not in ("mean", "sum", "none"):
            raise ValueError(f'Unsupported reduction: {reduction}, available options are ["mean", "sum", "none"].')
        self.sigmoid, self.squared_pred, self.reduction = sigmoid, squared_pred, reduction
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _DiceLossFn.apply(input, target, self.sigmoid, self.squared_pred, self.smooth_nr, self.smooth_dr, self.reduction)


def _metric_from_sums(s: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    inter, y_pred_o, y_o = s[..., 3], s[..., 4], s[..., 5]
    f = torch.where(y_o > 0, (2.0 * inter) / (y_o + y_pred_o), torch.full_like(inter, float("nan")))
    nans = torch.isnan(f)
    not_nans = (~nans).to(torch.float32)
    f = torch.where(nans, torch.zeros_like(f), f).to(torch.float32)
    zero = torch.zeros(1, device=f.device, dtype=f.dtype)
    not_nans = not_nans.sum(dim=1)
    f = torch.where(not_nans > 0, f.sum(dim=1) / not_nans, zero)
    not_nans = (not_nans > 0).float().sum(dim=0)
    f = torch.where(not_nans > 0, f.sum(dim=0) / not_nans, zero)
    return f, not_nans


class DiceMetric:
    """monai.metrics.DiceMetric(include_background=True, reduction="mean"): __call__(y_pred, y) with binarised
    predictions returns (mean Dice, not_nans).  `from_logits(logits, y)` fuses the scripts' Activations(sigmoid) ->
    AsDiscrete(threshold) in front of it (sigmoid(x) >= 0.5 <=> x >= 0)."""

    def __init__(self, include_background: bool = True, reduction: str = "mean") -> None:
        if not include_background or reduction != "mean":
            raise NotImplementedError("mvtb DiceMetric covers DiceMetric(include_background=True, reduction='mean')")

    def __call__(self, y_pred: torch.Tensor, y: torch.Tensor):
        return _metric_from_sums(dice_sums(y_pred, y, from_logits=False))

    def from_logits(self, logits: torch.Tensor, y: torch.Tensor):
        return _metric_from_sums(dice_sums(logits, y, from_logits=True))


def dice_loss_and_metric(logits: torch.Tensor, target: torch.Tensor, smooth_nr: float = 1e-5, smooth_dr: float = 1e-5):
    """Both from ONE pass (no gradient): (DiceLoss(sigmoid, squared_pred) value, (metric, not_nans))."""
    s = dice_sums(logits, target, from_logits=True)
    f = 1.0 - (2.0 * s[..., 0] + smooth_nr) / (s[..., 1] + s[..., 2] + smooth_dr)
    return f.mean().to(torch.float32), _metric_from_sums(s)


class _FreqConsistencyFn(torch.autograd.Function):
    """MSE(Re fftn real, Re fftn fake) + MSE(Im fftn real, Im fftn fake) over the last two axes (reconGan_freq.py:134-140)
    = H W * MSE(real, fake) by Parseval (the transform is unnormalised): one fused squared-difference reduction."""

    @staticmethod
    def forward(ctx, real, fake):
        require_cuda()
        if real.shape != fake.shape or real.dim() < 2:
            raise ValueError("freq_consistency_loss: shapes differ")
        a = real.detach().to(torch.float32).contiguous()
        b = fake.detach().to(torch.float32).contiguous()
        L = _lib.lib()
        out = torch.empty(1, dtype=torch.float64, device=a.device)
        scratch = torch.empty(max(int(L.mvtb_dice_scratch_bytes(1)), 8), dtype=torch.uint8, device=a.device)
        with torch.cuda.device(a.device):
            rc = L.mvtb_sqdiff_sum_f32(_ptr(a), _ptr(b), a.numel(), _ptr(out), _ptr(scratch), _stream(a.device))
        _lib.check(L, rc)
        hw = float(a.shape[-1] * a.shape[-2])
        ctx.save_for_backward(a, b)
        ctx.c = 2.0 * hw / max(a.numel(), 1)
        return (out[0] * (hw / max(a.numel(), 1))).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        d = (a - b) * (g * ctx.c)
        return d, -d


def freq_consistency_loss(real: torch.Tensor, fake: torch.Tensor) -> torch.Tensor:
    return _FreqConsistencyFn.apply(real, fake)
