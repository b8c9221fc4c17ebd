"""Drop-in for the reference's source_code/stylization_layers.py (S below): the artifact layers
that run on the GPU as the first layer of the segmentation network.

GibbsNoiseLayer.forward and spike_layer.forward are one fused k-space pass in libmvtb.so instead of
cuFFT C2C + ~15 elementwise kernels; names, attributes (`alpha`, `device`, `intensity`) and the
"n_dims = rank - 1" rule (a 5-D batch is a 4-D FFT over (C,H,W,D), S:81) are the reference's.
"""
import numpy as np
import torch
import torch.nn as nn

from filters_and_operators import RandKSpaceSpikeNoise
from mvtb import _lib, functional as Fn, host

# monai.networks.nets.UNet when MONAI is installed, else a plain-torch residual U-Net with the same constructor
# (mvtb/_monai_compat.py): the network consumes the layers' output and is not part of the hot path
from mvtb._monai_compat import UNet


class Fourier:
    """Centred k-space helpers kept for API compatibility (S:16-52); unused by the layers."""

    @staticmethod
    def shift_fourier(x: torch.Tensor, n_dims: int) -> torch.Tensor:
        axes = tuple(range(-n_dims, 0))
        return torch.fft.fftshift(torch.fft.fftn(x, dim=axes), dim=axes)

    @staticmethod
    def inv_shift_fourier(k: torch.Tensor, n_dims: int) -> torch.Tensor:
        axes = tuple(range(-n_dims, 0))
        return torch.fft.ifftn(torch.fft.ifftshift(k, dim=axes), dim=axes).real


class _MaskedSpectrum(torch.autograd.Function):
    """y = Re ifftn(M fftn(x)).  M_eff = (M(f)+M(-f))/2 is real and even, so the operator is
    self-adjoint: the input gradient is the same kernel applied to grad_output.  (d y / d alpha is
    identically zero for the hard mask, as in the reference - SURVEY A.9.)"""

    @staticmethod
    def forward(ctx, x, n_dims, thresh):
        ctx.n_dims, ctx.thresh = n_dims, thresh
        desc = host.make_desc(mask_kind=_lib.MASK_CENTRED, mask_ndim=n_dims, mask_thresh=thresh)
        return Fn.kspace_chain(x.contiguous(), n_dims, [desc])

    @staticmethod
    def backward(ctx, g):
        desc = host.make_desc(mask_kind=_lib.MASK_CENTRED, mask_ndim=ctx.n_dims, mask_thresh=ctx.thresh)
        return Fn.kspace_chain(g.contiguous(), ctx.n_dims, [desc]), None, None


class GibbsNoiseLayer(nn.Module, Fourier):
    """Gibbs noise layer (S:55-116): keep k-space where dist/(alpha*dist.max()) <= 1 around the (N-1)/2
    centre over k.shape[1:]; alpha = 1 is the identity.  `alpha` is a plain tensor attribute that
    callers reassign (finite-difference updates in 350_stylized_layers), not a Parameter."""

    def __init__(self, alpha=None) -> None:
        nn.Module.__init__(self)
        self.device = torch.device('cuda:0') if torch.cuda.is_available() else torch.device('cpu')
        if alpha is None:
            self.alpha = torch.rand(1, requires_grad=True, device=self.device)
        else:
            alpha = min(max(alpha, 0.), 1.)
            self.alpha = torch.tensor([alpha], requires_grad=True, device=self.device)

    def _alpha_on_host(self) -> float:
        """alpha as a Python float.  The mask threshold is computed on the host (S:71 reads alpha there too), which for a CUDA
        tensor costs a device-to-host copy and a stream synchronisation per forward; the value is read again only when
        the attribute was reassigned (what the 350_stylized_layers scripts do) or written in place (version counter; a write
        through `.data` bypasses that counter and is not seen), so that a loop that leaves alpha alone between
        finite-difference updates pays for it once per update."""
        a = self.alpha
        if not isinstance(a, torch.Tensor):
            return float(a)
        cached = self.__dict__.get("_alpha_cache")
        # the cache keeps the tensor itself (not its id(), which Python reuses once a tensor is freed)
        if cached is None or cached[0] is not a or cached[1] != a._version:
            cached = (a, a._version, float(a.detach().reshape(-1)[0]))
            self.__dict__["_alpha_cache"] = cached
        return cached[2]

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        n_dims = len(img.shape[1:])
        if n_dims < 2 or n_dims > 4:
            raise ValueError(f"GibbsNoiseLayer supports inputs of rank 3 to 5, got rank {img.dim()}")
        x, src = Fn.to_device(img)
        alpha = self._alpha_on_host()
        thresh = host.layer_threshold(np.float32(alpha), img.shape[1:])
        if x.requires_grad:
            y = _MaskedSpectrum.apply(x, n_dims, thresh)
        else:
            y = Fn.kspace_chain(x, n_dims, [host.make_desc(mask_kind=_lib.MASK_CENTRED, mask_ndim=n_dims, mask_thresh=thresh)])
        return Fn.back(y, src)


class Gibbs_UNet(nn.Module):
    """GibbsNoiseLayer(0.5) in front of the 3-D ResUNet (S:119-139); the argument is ignored, as in the reference."""

    def __init__(self, alpha=None):
        super().__init__()
        self.gibbs = GibbsNoiseLayer(.5)
        self.ResUnet = UNet(dimensions=3, in_channels=1, out_channels=1, channels=(16, 32, 64, 128, 256),
                            strides=(2, 2, 2, 2), num_res_units=2)

    def forward(self, img):
        return self.ResUnet(self.gibbs(img))


class spike_layer(nn.Module):
    """A fresh RandKSpaceSpikeNoise(prob=1, range=(I,I), channel_wise=False) per forward (S:143-151):
    one location over img.shape[1:] shared by the whole batch, drawn from the class-level stream."""

    def __init__(self, intensity):
        super().__init__()
        self.intensity = torch.tensor(intensity)

    def forward(self, x):
        i = self.intensity.item()
        return RandKSpaceSpikeNoise(prob=1., intensity_range=(i, i), channel_wise=False)(x)


class Spikes_UNet(nn.Module):
    """spike_layer in front of MONAI's 3-D ResUNet (S:154-174)."""

    def __init__(self, intensity=15):
        super().__init__()
        self.spike = spike_layer(intensity)
        self.ResUnet = UNet(dimensions=3, in_channels=1, out_channels=1, channels=(16, 32, 64, 128, 256),
                            strides=(2, 2, 2, 2), num_res_units=2)

    def forward(self, img):
        return self.ResUnet(self.spike(img))
