"""Drop-in for the hot-path pieces of the reference's 50_reconstruction/reconGan/utils2.py (U below): RandZF, the random
k-space zero-filling that degrades the reconstruction GAN's inputs (U:34-74), on the CUDA kernels of libmvtb.so.

RandZF keeps the reference's constructor, attribute (`p`), clamping warning and random draw: the mask is
`torch.rand(k.size())` from torch's global CPU generator (U:71), compared with `<= p`; the transform itself is one fused
k-space pass (mask kind MVTB_MASK_UNIFORM) instead of fftn + fftshift + indexed assignment + ifftshift + ifftn.
`rng="philox"` draws the field on the GPU instead (counter-based, seed / offset)."""
import warnings

import torch

from mvtb import _lib, functional as Fn, host
from mvtb._monai_compat import Transform


class FourierTransform:
    """Centred k-space helpers kept for API compatibility (U:6-31); RandZF does not use them."""

    @staticmethod
    def shift_fourier(x: torch.Tensor, n_dims: int) -> torch.Tensor:
        axes = tuple(range(-n_dims, 0))
        return torch.fft.fftshift(torch.fft.fftn(x, dim=axes), dim=axes)

    @staticmethod
    def inv_shift_fourier(k: torch.Tensor, n_dims: int) -> torch.Tensor:
        axes = tuple(range(-n_dims, 0))
        return torch.fft.ifftn(torch.fft.ifftshift(k, dim=axes), dim=axes).real


class RandZF(Transform, FourierTransform):
    """Random zero-filling in k-space (U:34-74): every k-space sample whose uniform draw is <= p is set to zero, the
    image is the real part of the inverse transform.  p = 0 is the identity."""

    def __init__(self, p: float = 0, rng: str = "torch", seed: int = 0):
        self.p = min(max(0, p), 1.)
        if p < 0 or p > 1:
            warnings.warn(f'Setting p to {self.p}.')
        if rng not in ("torch", "philox"):
            raise ValueError("rng must be 'torch' or 'philox'")
        self.rng, self.seed, self.offset = rng, int(seed), 0

    def __call__(self, img: torch.Tensor, u: torch.Tensor = None):
        """img: (C, spatial...) with 2 to 4 spatial axes.  u: optional injected uniform field of img's shape."""
        n_dims = len(img.size()[1:])
        if n_dims < 2 or n_dims > 4:
            raise ValueError(f"RandZF supports 2 to 4 spatial axes, got {n_dims}")
        x, org = Fn.to_device(img)
        if u is None:
            if self.rng == "torch":
                u = torch.rand(img.size())                                   # the reference's draw (U:71)
            else:
                u = Fn.philox_uniform(x.numel(), self.seed, self.offset, x.device).reshape(x.shape)
                self.offset += (x.numel() + 3) // 4
        ud = u.to(device=x.device, dtype=torch.float32).contiguous()
        n_vox = int(x[0].numel())
        descs = [host.make_desc(mask_kind=_lib.MASK_UNIFORM, mask_ndim=n_dims, mask_u=ud.data_ptr() + 4 * n_vox * c, mask_p=float(self.p))
                 for c in range(x.shape[0])]
        # keep -> u > p  (the reference zeroes where u <= p)
        return Fn.back(Fn.kspace_chain(x, n_dims, descs), org)

    def rand_mask(self, k: torch.Tensor):
        """The reference's helper on an explicit (complex) k-space tensor (U:63-74); plain torch, not a hot path."""
        mask = torch.rand(k.size())
        k = k.clone()
        k[mask.to(k.device) <= self.p] = 0
        return k


def weights_init(m):
    """DCGAN initialisation used by the reconstruction GAN (U:77-84)."""
    import torch.nn as nn
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
