"""Sparse (geometric-gap) salt-and-pepper on the GPU: bit-exact against the numpy restatement of the
sampler, Bernoulli(p) statistics at full size, and the values the reference assigns (min/2, max/2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p", [0.0, 0.05, 0.35, 1.0])
def test_bit_exact_against_restatement(cuda_device, p):
    from mvtb import functional as Fn
    from oracle import philox_ref as R, ref_port as P
    x = P.synthetic_volume(5, (3, 9, 11, 13))
    n_per = x[0].numel()
    xd = x.to(cuda_device)
    y = Fn.salt_pepper(xd, p, seed=77, offset=5, n_samples=3, sparse=True)
    assert y.data_ptr() != xd.data_ptr() and torch.equal(xd.cpu(), x)          # out-of-place call leaves x alone
    pos, kind = R.sparse_hits(n_per, 3, 77, 5, p)
    want = x.clone().reshape(-1)
    smp = torch.from_numpy(pos // n_per)
    lo = torch.stack([x[s].min() / 2 for s in range(3)])
    hi = torch.stack([x[s].max() / 2 for s in range(3)])
    want[torch.from_numpy(pos)] = torch.where(torch.from_numpy(kind) == 1, hi[smp], lo[smp])
    assert torch.equal(y.cpu().reshape(-1), want)
    z = xd.clone()
    assert Fn.salt_pepper(z, p, seed=77, offset=5, n_samples=3, sparse=True, out=z).data_ptr() == z.data_ptr()
    assert torch.equal(z, y)


def test_statistics_full_size(cuda_device):
    from mvtb import functional as Fn
    from oracle import ref_port as P
    x = (P.synthetic_volume(3, (2, 1, 240, 240, 155)) + 0.01).to(cuda_device)
    n = x[0].numel()
    for p in (0.05, 0.15):
        y = Fn.salt_pepper(x, p, seed=11, offset=0, n_samples=2, sparse=True)
        for s in range(2):
            hit = y[s] != x[s]
            lo, hi = x[s].min() / 2, x[s].max() / 2
            vals = y[s][hit]
            assert bool(((vals == lo) | (vals == hi)).all())
            frac = float(hit.float().mean())
            assert abs(frac - p) < 5 * np.sqrt(p * (1 - p) / n)
            assert abs(float((vals == hi).float().mean()) - 0.5) < 5 * np.sqrt(0.25 / int(hit.sum()))
        assert not torch.equal(y[0] != x[0], y[1] != x[1])                       # samples get different fields


def test_drop_in_class_sparse_mode(cuda_device):
    import filters_and_operators as F
    x = torch.randn(1, 64, 48, 40, device=cuda_device)
    tr = F.SaltAndPepper(0.2, rng="philox-sparse", seed=3)
    a, b = tr.salt_and_pepper(x), tr.salt_and_pepper(x)
    fa = float((a != x).float().mean())
    assert 0.17 < fa < 0.23 and not torch.equal(a, b)
