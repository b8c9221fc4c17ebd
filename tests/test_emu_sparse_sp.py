"""Sparse (geometric-gap) salt-and-pepper through the DEBUG emulator: bit-exact against the numpy
restatement of the sampler (oracle/philox_ref.py) and statistically Bernoulli(p)."""
import ctypes as C
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")

from cuemu import emu  # noqa: E402
from mvtb import _lib as B  # noqa: E402
from oracle import philox_ref as R, ref_port as P  # noqa: E402


def run(x, n_samples, seed, offset, p, mm):
    L = emu.lib()
    z = x.copy()
    tab = np.zeros(B.SP_BLOCK, dtype=np.uint32)
    B.check(L, L.mvtb_salt_pepper_sparse_f32(emu.ptr(z), z.size // n_samples, n_samples, seed, offset, C.c_float(p),
                                             emu.ptr(mm), emu.ptr(tab), None))
    return z


def test_table_matches_restatement():
    for p in (0.0, 0.05, 0.15, 0.35, 1.0):
        t = (C.c_uint32 * B.SP_BLOCK)()
        B.check(emu.lib(), emu.lib().mvtb_sparse_table(C.c_float(p), t))
        assert np.array_equal(np.array(t), R.sparse_table(p))


@pytest.mark.parametrize("p", [0.0, 0.05, 0.3, 1.0])
def test_bit_exact_against_restatement(p):
    x = P.synthetic_volume(5, (3, 9, 11, 13)).numpy()            # 3 samples of 1287 voxels: partial last block
    n_per = x[0].size
    mm = np.zeros(6, dtype=np.float32)
    B.check(emu.lib(), emu.lib().mvtb_minmax_f32(emu.ptr(x), n_per, 3, emu.ptr(mm), None))
    z = run(x, 3, 77, 5, p, mm)
    pos, kind = R.sparse_hits(n_per, 3, 77, 5, p)
    want = x.copy().reshape(-1)
    smp = pos // n_per
    want[pos] = np.where(kind == 1, 0.5 * mm[2 * smp + 1], 0.5 * mm[2 * smp])
    assert np.array_equal(z.reshape(-1), want)


def test_statistics():
    n = 1 << 20
    x = np.zeros((1, n), dtype=np.float32)
    mm = np.array([-2.0, 4.0], dtype=np.float32)
    for p in (0.05, 0.25):
        z = run(x, 1, 2024, 0, p, mm).reshape(-1)
        hit = z != 0
        frac = hit.mean()
        assert abs(frac - p) < 4 * np.sqrt(p * (1 - p) / n)
        salt = (z[hit] == 2.0).mean()
        assert abs(salt - 0.5) < 4 * np.sqrt(0.25 / hit.sum())
        assert set(np.unique(z)) <= {0.0, -1.0, 2.0}
        # no structure at the 256-voxel block period: hits per position-in-block are flat
        per_pos = hit.reshape(-1, 256).mean(axis=0)
        assert abs(per_pos - p).max() < 6 * np.sqrt(p * (1 - p) / (n / 256))
        # different offset -> different field
        assert not np.array_equal(z, run(x, 1, 2024, 4096, p, mm).reshape(-1))
