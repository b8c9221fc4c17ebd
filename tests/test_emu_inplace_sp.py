"""In-place salt-and-pepper (the sparse variant that never reads x) through the DEBUG emulator."""
import ctypes as C
import shutil

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")

from cuemu import emu  # noqa: E402
from mvtb import _lib as B  # noqa: E402
from oracle import philox_ref, ref_port as P  # noqa: E402


@pytest.mark.parametrize("philox", [False, True])
def test_inplace_matches_out_of_place_and_oracle(philox):
    L = emu.lib()
    x = P.synthetic_volume(5, (3, 8, 5, 4)).numpy()           # 3 samples of 160 voxels (multiple of 4: vector path)
    n_per = x[0].size
    mm = np.zeros(6, dtype=np.float32)
    B.check(L, L.mvtb_minmax_f32(emu.ptr(x), n_per, 3, emu.ptr(mm), None))
    if philox:
        u, uarg = philox_ref.uniform_f32(x.size, 99, 3).reshape(x.shape), None
    else:
        u = np.stack([P.synthetic_uniform(i, x.shape[1:]).numpy() for i in range(3)])
        uarg = emu.ptr(u)
    y = np.empty_like(x)
    B.check(L, L.mvtb_salt_pepper_f32(emu.ptr(x), emu.ptr(y), n_per, 3, uarg, 99, 3, C.c_float(0.3), emu.ptr(mm), None))
    z = x.copy()
    B.check(L, L.mvtb_salt_pepper_f32(emu.ptr(z), emu.ptr(z), n_per, 3, uarg, 99, 3, C.c_float(0.3), emu.ptr(mm), None))
    assert np.array_equal(y, z)
    for s in range(3):
        want = P.salt_and_pepper(torch.from_numpy(x[s]), 0.3, torch.from_numpy(np.ascontiguousarray(u[s]))).numpy()
        assert np.array_equal(z[s], want)
