"""Kernel sources run through the DEBUG emulator (tests/cuemu) against the golden vectors.

These tests exist because the build container has no GPU: they execute the same .cu sources,
compiled by g++ against a fiber-based CUDA shim, to check index arithmetic and the C-ABI
argument handling.  They do not count as parity evidence for the CUDA path; the `-m gpu`
tests (tests/test_gpu_*.py) are the parity tests proper and go through libmvtb.so.
"""
import ctypes as C
import shutil

import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")

from cuemu import emu  # noqa: E402
from mvtb import _lib as B, host  # noqa: E402
from oracle import ref_port as P  # noqa: E402

TOL = 1e-5   # north_star: relative L2 <= 1e-5 for FFT-based outputs


def pick(names, k):
    return names[:: max(1, len(names) // k)][:k]


@pytest.mark.parametrize("name", golden_names("disk_"))
def test_emu_disk(name):
    m, z = load_golden(name)
    r = float("inf") if m["r"] == "inf" else m["r"]
    x = z["x"]
    d = host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=host.disk_threshold(r, x.shape[-3:]),
                       inside_off=m["inside_off"])
    y, _ = emu.chain(x, 3, [d])
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("gibbs_"))
def test_emu_gibbs(name):
    m, z = load_golden(name)
    x = z["x"]
    nd = x.ndim - 1
    d = host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=nd, mask_thresh=host.gibbs_threshold(m["alpha"], x.shape[1:]))
    y, _ = emu.chain(x, nd, [d])
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("layer_"))
def test_emu_layer(name):
    m, z = load_golden(name)
    x = z["x"]
    nd = x.ndim - 1
    a = np.float32(min(max(m["alpha"], 0.), 1.))
    d = host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=nd, mask_thresh=host.layer_threshold(a, x.shape[1:]))
    y, _ = emu.chain(x, nd, [d])
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("wrap_"))
def test_emu_wrap_chain(name):
    m, z = load_golden(name)
    d = host.make_desc(wrap_alpha=m["alpha"], wrap_naxes=3)
    y, _ = emu.chain(z["x"], 3, [d])
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("planes_"))
def test_emu_planes(name):
    m, z = load_golden(name)
    d = host.make_desc(spikes=[(m["idx"], host.exp_f32(m["intensity"]))])
    y, _ = emu.chain(z["x"], 3, [d])
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("kspike_"))
def test_emu_kspike(name):
    m, z = load_golden(name)
    x = z["x"]
    nd = x.ndim - 1
    C_ = x.shape[0]
    loc, inten = m["loc"], m["intensity"]
    locs = loc if isinstance(loc[0], list) else [loc]
    if inten is None:
        ints = [float(v) for v in z["logabs_mean25"]]          # F:932-933: per-channel default, zipped with locs (F:937)
        ints = [ints[l[0]] for l in locs] if not isinstance(loc[0], list) else ints[:len(locs)]
    else:
        ints = inten if isinstance(inten, list) else [inten]
    per_chan = [[] for _ in range(C_)]
    for l, v in zip(locs, ints):
        if len(l) == x.ndim:
            per_chan[l[0]].append((l[1:], host.exp_f32(v)))
        else:
            for c in range(C_):
                per_chan[c].append((l, host.exp_f32(v)))
    descs = [host.make_desc(spikes=s) for s in per_chan]
    y, _ = emu.chain(x, nd, descs)
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", golden_names("kspike_"))
def test_emu_logabs_mean(name):
    m, z = load_golden(name)
    x = np.ascontiguousarray(z["x"])
    nd = x.ndim - 1
    plan = emu.Plan(x.shape[1:], 1)
    sums = np.zeros(x.shape[0], dtype=np.float64)
    B.check(emu.lib(), emu.lib().mvtb_kspace_logabs_sum_f32(plan.h, emu.ptr(x), x.shape[0], emu.ptr(sums), None))
    got = 2.5 * sums / np.prod(x.shape[1:])
    assert np.allclose(got, z["logabs_mean25"], rtol=2e-6, atol=1e-5)


@pytest.mark.parametrize("name", golden_names("chain127_"))
def test_emu_chain127(name):
    m, z = load_golden(name)
    x = z["x"]
    thr = host.disk_threshold(m["r"], x.shape[-3:])
    # stage-wise, each stage fed the reference's previous-stage output
    y1, _ = emu.chain(x, 3, [host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=thr)])
    assert rel_l2(y1, z["y1"]) <= TOL
    y3, _ = emu.chain(z["y2"], 3, [host.make_desc(wrap_alpha=m["alpha"])])
    assert rel_l2(y3, z["y3"]) <= TOL
    amp = host.exp_f32(m["intensity"])
    if m["spike_in_ball"]:
        y2, _ = emu.chain(z["y1"], 3, [host.make_desc(spikes=[(m["idx"], amp)])])
        assert rel_l2(y2, z["y2"]) <= TOL
        # the whole k-space part in ONE pass, with per-sample min/max for the S&P that follows
        d = host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=thr, spikes=[(m["idx"], amp)], wrap_alpha=m["alpha"])
        yf, mm = emu.chain(x, 3, [d], minmax_vols_per_sample=x.shape[0])
        assert rel_l2(yf, z["y3"]) <= TOL
        assert np.allclose(mm, [z["y3"].min(), z["y3"].max()], rtol=1e-4)


@pytest.mark.parametrize("shape", [(1, 8, 6, 37), (2, 41, 4, 6), (1, 6, 74, 5), (1, 43, 37, 47), (2, 82, 9)])
def test_emu_prime_factors_above_31(shape):
    """Axis lengths with prime factors > 31 run through the generic direct-DFT stage (any length works)."""
    x = P.synthetic_volume(9, shape).numpy() + 0.1
    nd = len(shape) - 1
    d = host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=nd, mask_thresh=host.gibbs_threshold(0.4, shape[1:]))
    y, _ = emu.chain(x, nd, [d])
    assert rel_l2(y, P.gibbs_noise(torch.from_numpy(x), 0.4).numpy()) <= TOL
    y, _ = emu.chain(x, nd, [host.make_desc()])                  # plain round trip
    assert rel_l2(y, x) <= TOL


def test_emu_chunking_and_per_volume_descs():
    """n_volumes > chunk, distinct desc per volume, odd row count (zero-padded pair), min/max per sample."""
    x = P.synthetic_volume(21, (5, 3, 5, 6)).numpy()          # 5 volumes of 3x5x6: 15 rows each
    descs = []
    want = []
    for c in range(5):
        r = 1.5 + 0.5 * c
        descs.append(host.make_desc(mask_kind=B.MASK_DISK, mask_ndim=3, mask_thresh=host.disk_threshold(r, (3, 5, 6))))
        want.append(P.fourier_disk_mask(torch.from_numpy(x[c:c + 1]), r).numpy())
    want = np.concatenate(want)
    y, mm = emu.chain(x, 3, descs, chunk=2, minmax_vols_per_sample=2)
    assert rel_l2(y, want) <= TOL
    for s in range(3):
        blk = y[2 * s: 2 * s + 2]
        assert mm[2 * s] == blk.min() and mm[2 * s + 1] == blk.max()


def test_emu_general_path_many_per_volume_descs_and_identical_descs():
    """General FFT path (GibbsNoise mask: not band-limited) with one descriptor per volume for more volumes than a
    chunk holds, so descriptors are indexed by volume across chunks; and n identical descriptors == one descriptor."""
    shape = (11, 6, 5, 9)                                     # 11 volumes of 6x5x9
    x = P.synthetic_volume(22, shape).numpy()
    alphas = [0.15 + 0.07 * c for c in range(11)]
    descs = [host.make_desc(mask_kind=B.MASK_CENTRED, mask_ndim=3, mask_thresh=host.gibbs_threshold(a, shape[1:])) for a in alphas]
    y, _ = emu.chain(x, 3, descs, chunk=4)
    for c in range(11):
        want = P.gibbs_noise(torch.from_numpy(x[c:c + 1]), alphas[c]).numpy()
        assert rel_l2(y[c:c + 1], want) <= TOL, c
    y1, _ = emu.chain(x, 3, [descs[3]], chunk=4)
    yn, _ = emu.chain(x, 3, [descs[3]] * 11, chunk=4)
    assert np.array_equal(y1, yn)


@pytest.mark.parametrize("name", golden_names("sap_"))
def test_emu_salt_pepper_injected_uniforms_bit_exact(name):
    m, z = load_golden(name)
    x, u = np.ascontiguousarray(z["x"]), np.ascontiguousarray(z["u"])
    mm = np.zeros(2, dtype=np.float32)
    L = emu.lib()
    B.check(L, L.mvtb_minmax_f32(emu.ptr(x), x.size, 1, emu.ptr(mm), None))
    assert mm[0] == x.min() and mm[1] == x.max()
    y = np.empty_like(x)
    B.check(L, L.mvtb_salt_pepper_f32(emu.ptr(x), emu.ptr(y), x.size, 1, emu.ptr(u), 0, 0, C.c_float(m["p"]), emu.ptr(mm), None))
    assert np.array_equal(y, z["y"])


def test_emu_salt_pepper_philox_matches_oracle_given_same_uniforms():
    from oracle import philox_ref
    L = emu.lib()
    x = P.synthetic_volume(5, (3, 7, 5, 9)).numpy()           # 3 samples of 315 voxels (not a multiple of 4)
    n_per = x[0].size
    u = np.empty(x.size, dtype=np.float32)
    B.check(L, L.mvtb_philox_uniform_f32(emu.ptr(u), u.size, 1234567890123, 77, None))
    assert np.array_equal(u, philox_ref.uniform_f32(u.size, 1234567890123, 77))
    mm = np.zeros(6, dtype=np.float32)
    B.check(L, L.mvtb_minmax_f32(emu.ptr(x), n_per, 3, emu.ptr(mm), None))
    y = np.empty_like(x)
    B.check(L, L.mvtb_salt_pepper_f32(emu.ptr(x), emu.ptr(y), n_per, 3, None, 1234567890123, 77, C.c_float(0.3), emu.ptr(mm), None))
    uu = u.reshape(x.shape)
    for s in range(3):
        want = P.salt_and_pepper(torch.from_numpy(x[s]), 0.3, torch.from_numpy(uu[s])).numpy()
        assert np.array_equal(y[s], want)


@pytest.mark.parametrize("name", [n for n in golden_names("wrap_") if "s9x15x25" not in n and "155" not in n and "x31" not in n])
def test_emu_wrap_fold(name):
    m, z = load_golden(name)
    x = np.ascontiguousarray(z["x"])
    y = np.empty_like(x)
    L = emu.lib()
    rc = L.mvtb_wrap_fold_f32(emu.ptr(x), emu.ptr(y), x.shape[0], x.shape[1], x.shape[2], x.shape[3], C.c_float(m["alpha"]), None)
    B.check(L, rc)
    assert rel_l2(y, z["y"]) <= TOL


@pytest.mark.parametrize("name", [n for n in golden_names("wrap_") if "s9x15x25" not in n])
def test_emu_wrap_odd_last(name):
    """Even H and W, any D (incl. 31, 155 and the even ones): folds along H, W + one-kernel FFT filter along D."""
    m, z = load_golden(name)
    x = np.ascontiguousarray(z["x"])
    y = np.empty_like(x)
    L = emu.lib()
    plan = emu.Plan(x.shape[1:], 2)
    B.check(L, L.mvtb_wrap_odd_last_f32(plan.h, emu.ptr(x), emu.ptr(y), x.shape[0], C.c_float(m["alpha"]), None))
    assert rel_l2(y, z["y"]) <= TOL


def test_emu_wrap_odd_last_rejects_odd_h_or_w():
    x = np.zeros((1, 9, 15, 25), dtype=np.float32)
    L = emu.lib()
    plan = emu.Plan(x.shape[1:], 1)
    assert L.mvtb_wrap_odd_last_f32(plan.h, emu.ptr(x), emu.ptr(np.empty_like(x)), 1, C.c_float(0.5), None) == B.MVTB_EUNSUPPORTED


def test_emu_wrap_fold_rejects_odd_axis():
    x = np.zeros((1, 4, 6, 5), dtype=np.float32)
    L = emu.lib()
    assert L.mvtb_wrap_fold_f32(emu.ptr(x), emu.ptr(np.empty_like(x)), 1, 4, 6, 5, C.c_float(0.5), None) == B.MVTB_EUNSUPPORTED
    assert "odd axis" in B.last_error(L)


def test_emu_argument_errors():
    L = emu.lib()
    h = C.c_void_p()
    shp = (C.c_int * 3)(8, 8, 40009)       # a prime axis whose tile (plus scratch) cannot fit in shared memory
    assert L.mvtb_plan_create(C.byref(h), 3, shp, 1, 0) == B.MVTB_EUNSUPPORTED
    assert L.mvtb_plan_create(C.byref(h), 5, shp, 1, 0) == B.MVTB_EINVAL
    plan = emu.Plan((4, 6, 8))
    x = np.zeros((1, 4, 6, 8), dtype=np.float32)
    d = host.make_desc(spikes=[((4, 0, 0), 1.0)])      # index 4 out of bounds on an axis of length 4
    assert L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(x), 1, host.desc_array([d]), 1, None, 1, None) == B.MVTB_EINVAL
    assert L.mvtb_kspace_chain_f32(plan.h, emu.ptr(x), emu.ptr(x), 3, host.desc_array([d, d]), 2, None, 1, None) == B.MVTB_EINVAL
