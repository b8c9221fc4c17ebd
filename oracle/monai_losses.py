"""CPU restatement of the MONAI 0.5 Dice loss / metric the reference's scripts call after the hot path.  TEST
INFRASTRUCTURE ONLY.  (MONAI 0.5.dev2113 is not part of /root/reference and not installed here; the published
algorithm of monai/losses/dice.py:DiceLoss.forward and monai/metrics/meandice.py:compute_meandice +
monai/metrics/utils.py:do_metric_reduction is restated with the same torch ops.)

Call sites: 10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:216 (DiceLoss(to_onehot_y=False,
sigmoid=True, squared_pred=True)), :266-283 (Activations(sigmoid) -> AsDiscrete(threshold_values=True) -> DiceMetric(
include_background=True, reduction="mean")), 350_stylized_layers/gibbs0p7_layer_domain_GD.py:252-269."""
import torch


def dice_loss(logits: torch.Tensor, target: torch.Tensor, sigmoid: bool = True, squared_pred: bool = True,
              smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, reduction: str = "mean") -> torch.Tensor:
    inp = torch.sigmoid(logits) if sigmoid else logits
    reduce_axis = list(range(2, inp.dim()))
    intersection = torch.sum(target * inp, dim=reduce_axis)
    if squared_pred:
        target = torch.pow(target, 2)
        inp = torch.pow(inp, 2)
    ground_o = torch.sum(target, dim=reduce_axis)
    pred_o = torch.sum(inp, dim=reduce_axis)
    denominator = ground_o + pred_o
    f = 1.0 - (2.0 * intersection + smooth_nr) / (denominator + smooth_dr)
    if reduction == "mean":
        return torch.mean(f)
    if reduction == "sum":
        return torch.sum(f)
    return f


def post_trans(logits: torch.Tensor) -> torch.Tensor:
    """Activations(sigmoid=True) -> AsDiscrete(threshold_values=True) (logit_thresh 0.5)."""
    return (torch.sigmoid(logits) >= 0.5).float()


def dice_metric(y_pred: torch.Tensor, y: torch.Tensor):
    """DiceMetric(include_background=True, reduction="mean")(y_pred, y) -> (value, not_nans)."""
    y = y.float()
    y_pred = y_pred.float()
    reduce_axis = list(range(2, y_pred.dim()))
    intersection = torch.sum(y * y_pred, dim=reduce_axis)
    y_o = torch.sum(y, reduce_axis)
    y_pred_o = torch.sum(y_pred, dim=reduce_axis)
    denominator = y_o + y_pred_o
    f = torch.where(y_o > 0, (2.0 * intersection) / denominator, torch.tensor(float("nan"), device=y_o.device))
    nans = torch.isnan(f)
    not_nans = (~nans).float()
    f = f.clone()
    f[nans] = 0
    t_zero = torch.zeros(1, device=f.device, dtype=f.dtype)
    not_nans = not_nans.sum(dim=1)
    f = torch.where(not_nans > 0, f.sum(dim=1) / not_nans, t_zero)       # channel average
    not_nans = (not_nans > 0).float().sum(dim=0)
    f = torch.where(not_nans > 0, f.sum(dim=0) / not_nans, t_zero)       # batch average
    return f, not_nans
