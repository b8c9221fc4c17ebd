"""CPU restatement of the reconstruction-GAN pieces next to the hot path.  TEST INFRASTRUCTURE ONLY.
U = /root/reference/50_reconstruction/reconGan/utils2.py, G = .../reconGan_freq.py."""
import torch
from torch.fft import fftn, fftshift, ifftn, ifftshift


def rand_zf(img: torch.Tensor, p: float, u: torch.Tensor) -> torch.Tensor:
    """RandZF.__call__ with the uniform field injected (U:53-74): u stands for `torch.rand(k.size())` (U:71)."""
    p = min(max(0, p), 1.)
    n_dims = len(img.size()[1:])
    axes = tuple(range(-n_dims, 0))
    k = fftshift(fftn(img, dim=axes), dim=axes)
    k = k.clone()
    k[u <= p] = 0
    return ifftn(ifftshift(k, dim=axes), dim=axes).real


def freq_consistency(real_batch: torch.Tensor, fake: torch.Tensor) -> torch.Tensor:
    """G:134-140 with l2_loss = nn.MSELoss() (G:60)."""
    l2 = torch.nn.MSELoss()
    real_k = fftn(real_batch, dim=(-2, -1))
    fake_k = fftn(fake, dim=(-2, -1))
    return l2(real_k.real, fake_k.real) + l2(real_k.imag, fake_k.imag)
