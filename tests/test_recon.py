"""RandZF (random k-space zero-filling) and the frequency-consistency loss of the reconstruction
GAN.  The restatement against the reference's golden vectors (CPU), the chain's uniform-mask kind through the DEBUG
emulator, and the CUDA path against both (-m gpu)."""
import shutil

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2
from oracle import recon_port as RP


@pytest.mark.parametrize("name", golden_names("randzf_"))
def test_restatement_equals_reference_golden(name):
    m, z = load_golden(name)
    y = RP.rand_zf(torch.from_numpy(z["x"]), m["p"], torch.from_numpy(z["u"]))
    assert torch.equal(y, torch.from_numpy(z["y"]))


def test_freq_consistency_is_hw_times_mse():
    _, z = load_golden("freqloss_s4x1x16x12")
    a, b = torch.from_numpy(z["a"]), torch.from_numpy(z["b"])
    assert abs(float(RP.freq_consistency(a, b)) - float(z["loss"])) < 1e-9
    assert abs(float(z["loss"]) - 16 * 12 * float(torch.nn.functional.mse_loss(a.double(), b.double()))) < 1e-4 * float(z["loss"])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")
@pytest.mark.parametrize("name", golden_names("randzf_"))
def test_emulated_uniform_mask_kind(name):
    from cuemu import emu
    from mvtb import _lib as B, host
    m, z = load_golden(name)
    x, u = np.ascontiguousarray(z["x"]), np.ascontiguousarray(z["u"])
    nd = x.ndim - 1
    L = emu.lib()
    plan = emu.Plan(x.shape[1:], 2)
    nv = x.shape[0]
    n = int(np.prod(x.shape[1:]))
    descs = [host.make_desc(mask_kind=B.MASK_UNIFORM, mask_ndim=nd, mask_u=u.ctypes.data + 4 * n * c, mask_p=m["p"]) for c in range(nv)]
    y = np.empty_like(x)
    B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(y), nv, host.desc_array(descs), nv, None, 1, None))
    assert rel_l2(y, z["y"]) <= 1e-5 if np.abs(z["y"]).max() > 0 else np.abs(y).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_names("randzf_"))
def test_gpu_randzf_golden(name, cuda_device):
    import utils2 as U
    m, z = load_golden(name)
    x, u = torch.from_numpy(z["x"]), torch.from_numpy(z["u"])
    y = U.RandZF(m["p"])(x, u=u)
    assert y.shape == x.shape and y.dtype == torch.float32 and y.device == x.device
    if np.abs(z["y"]).max() > 0:
        assert rel_l2(y.numpy(), z["y"]) <= 1e-5
    else:
        assert float(y.abs().max()) < 1e-6
    torch.manual_seed(m["seed"])                        # the transform's own draw is the reference's draw
    y2 = U.RandZF(m["p"])(x)
    assert rel_l2(y2.numpy(), y.numpy()) <= 1e-6 if np.abs(z["y"]).max() > 0 else True


@pytest.mark.gpu
def test_gpu_randzf_script_shape_and_freq_loss(cuda_device):
    import utils2 as U
    from mvtb import losses as LS
    g = torch.Generator().manual_seed(3)
    x = torch.randn(16, 128, 128, generator=g)          # a stack of slices, one mask per slice
    u = torch.rand(x.size(), generator=g)
    y = U.RandZF(0.2)(x.to(cuda_device), u=u)
    ref = RP.rand_zf(x, 0.2, u)
    assert y.is_cuda and rel_l2(y.cpu().numpy(), ref.numpy()) <= 1e-5
    a = torch.randn(8, 1, 128, 128, generator=g).to(cuda_device)
    b = torch.randn(8, 1, 128, 128, generator=g).to(cuda_device).requires_grad_(True)
    loss = LS.freq_consistency_loss(a, b)
    loss.backward()
    br = b.detach().clone().requires_grad_(True)
    ref_l = RP.freq_consistency(a, br)
    ref_l.backward()
    assert abs(float(loss) - float(ref_l)) <= 1e-5 * float(ref_l)
    assert rel_l2(b.grad.cpu().numpy(), br.grad.cpu().numpy()) <= 1e-5
