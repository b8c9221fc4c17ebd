"""One Gibbs_GD finite-difference step of the 350_stylized_layers scripts (gibbs0p7_layer_domain_GD.py:252-269) on the
scripts' batch (2, 1, 128, 128, 64): two model forwards (GibbsNoiseLayer + 3-D ResUNet) + two DiceLoss(sigmoid,
squared_pred) evaluations + the alpha update, timed host-to-completion; and the part this library owns on its own (two
layer forwards + two fused Dice reductions).  The UNet is MONAI's when importable, else the plain-torch stand-in with its
constructor (mvtb/_monai_compat.py): its time is torch's, reported for context only."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "medical-vision-textural-bias_b200"))
import stylization_layers as S  # noqa: E402
from mvtb import losses  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
inputs = torch.randn(2, 1, 128, 128, 64, device=dev)
labels = (torch.rand(2, 1, 128, 128, 64, device=dev) > 0.7).float()
model = S.Gibbs_UNet(0.7).to(dev).eval()
model.gibbs.alpha = torch.tensor([0.7], device=dev)
loss_function = losses.DiceLoss(to_onehot_y=False, sigmoid=True, squared_pred=True)


@torch.no_grad()
def Gibbs_GD(net, h=0.01, learning_rate=0.02):
    old_alpha = model.gibbs.alpha.clone()
    loss_0 = loss_function(net(inputs), labels)
    model.gibbs.alpha = old_alpha + h
    loss_h = loss_function(net(inputs), labels)
    delta = (loss_h - loss_0) / h
    model.gibbs.alpha = old_alpha - learning_rate * delta
    return loss_0.detach().item(), model.gibbs.alpha.item()


def wall(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


full = wall(lambda: Gibbs_GD(model))
own = wall(lambda: Gibbs_GD(model.gibbs))                 # the layer's output taken as logits: layer forwards + Dice only
print(f"Gibbs_GD step, batch (2,1,128,128,64): {full:.2f} ms with the UNet forwards ({'MONAI' if S.UNet.__module__.startswith('monai') else 'stand-in'} UNet, torch), "
      f"{own:.3f} ms for the two GibbsNoiseLayer forwards + two fused Dice losses + alpha update (this library); alpha now {float(model.gibbs.alpha):.4f}")
