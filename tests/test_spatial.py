"""Crop + flip in front of the intensity prologue ("RandFlipd, crop"): the MONAI 0.5 restatement
against known answers, the kernel through the DEBUG emulator and on the GPU against the restatement (data movement: equal
bit for bit), and the drop-in classes against the restatement's draw order."""
import ctypes as C
import shutil

import numpy as np
import pytest
import torch

from oracle import monai_spatial as M, ref_port as P

emu_only = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the emulator build")


def test_restatement_known_answers():
    img = np.arange(2 * 6 * 5 * 4, dtype=np.float32).reshape(2, 6, 5, 4)
    # centre crop: start = max(N // 2 - roi // 2, 0)
    s = M.center_crop_slices((6, 5, 4), (4, 3, 2))
    assert s == (slice(1, 5), slice(1, 4), slice(1, 3))
    s = M.center_crop_slices((6, 5, 4), (8, -1, 3))              # larger than the image / fall back to the image size
    assert s == (slice(0, 6), slice(0, 5), slice(1, 4))
    # random crop: one randint(0, N - roi + 1) per axis larger than the roi, in axis order
    R = np.random.RandomState(7)
    s = M.rand_spatial_crop_slices((6, 5, 4), (4, 5, 2), R)
    R2 = np.random.RandomState(7)
    o0 = R2.randint(0, 3); o2 = R2.randint(0, 3)                  # axis 1 is not larger than the roi: no draw
    assert s == (slice(o0, o0 + 4), slice(0, 5), slice(o2, o2 + 2))
    # flip after crop
    out = M.crop_then_flip(img, s, True, 0)
    assert out.shape == (2, 4, 5, 2) and np.array_equal(out[:, 0], img[:, o0 + 3, :, o2:o2 + 2])
    out = M.crop_then_flip(img, s, True, None)
    assert np.array_equal(out[1, 0, 0, 0], img[1, o0 + 3, 4, o2 + 1])
    assert np.array_equal(M.crop_then_flip(img, s, False, 0), img[:, o0:o0 + 4, :, o2:o2 + 2])


def _cases():
    return [((2, 9, 11, 13), (5, 7, 6), (3, 2, 4), 0), ((1, 16, 12, 10), (16, 12, 10), (0, 0, 0), 1), ((3, 8, 9, 7), (4, 9, 3), (4, 0, 2), 5),
            ((2, 6, 5, 40), (3, 2, 37), (1, 3, 3), 7), ((1, 4, 4, 4), (1, 1, 1), (3, 3, 3), 2)]


def _want(x, size, start, mask):
    sl = tuple(slice(o, o + s) for o, s in zip(start, size))
    out = x[(slice(None),) + sl]
    for a in range(3):
        if mask >> a & 1:
            out = np.flip(out, a + 1)
    return np.array(out, dtype=x.dtype, order="C", copy=True)


@emu_only
@pytest.mark.parametrize("shape,size,start,mask", _cases())
def test_emulated_kernel_is_the_gather(shape, size, start, mask):
    from cuemu import emu
    from mvtb import _lib as B
    L = emu.lib()
    x = P.synthetic_volume(11, shape).numpy()
    out = np.full((shape[0],) + size, -7.0, dtype=np.float32)
    i3 = C.c_int32 * 3
    B.check(L, L.mvtb_crop_flip_f32(emu.ptr(x), emu.ptr(out), shape[0], i3(*shape[1:]), i3(*size), i3(*start), mask, None))
    assert np.array_equal(out, _want(x, size, start, mask))


@emu_only
def test_emulated_kernel_rejects_bad_windows():
    from cuemu import emu
    L = emu.lib()
    x = np.zeros((1, 4, 4, 4), dtype=np.float32)
    out = np.zeros((1, 2, 2, 2), dtype=np.float32)
    i3 = C.c_int32 * 3
    assert L.mvtb_crop_flip_f32(emu.ptr(x), emu.ptr(out), 1, i3(4, 4, 4), i3(2, 2, 2), i3(3, 0, 0), 0, None) == -1     # 3 + 2 > 4
    assert L.mvtb_crop_flip_f32(emu.ptr(x), emu.ptr(out), 1, i3(4, 4, 4), i3(2, 2, 2), i3(0, 0, 0), 8, None) == -1     # flip bit 3
    assert L.mvtb_crop_flip_f32(emu.ptr(x), emu.ptr(x), 1, i3(4, 4, 4), i3(4, 4, 4), i3(0, 0, 0), 0, None) == -1       # in place


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size,start,mask", _cases() + [((4, 240, 240, 155), (128, 128, 64), (57, 101, 33), 1)])
def test_gpu_kernel_is_the_gather(cuda_device, shape, size, start, mask):
    from mvtb import spatial as S
    x = P.synthetic_volume(12, shape)
    got = S.crop_flip(x.to(cuda_device), start, size, mask)
    assert got.is_cuda and torch.equal(got.cpu(), torch.from_numpy(_want(x.numpy(), size, start, mask)))
    got_host = S.crop_flip(x.numpy(), start, size, mask)            # host input: result comes back to the host
    assert not got_host.is_cuda and torch.equal(got_host, got.cpu())


@pytest.mark.gpu
def test_gpu_dropins_follow_monai_draw_order(cuda_device):
    """RandSpatialCropd -> RandFlipd on image and label with Compose-style seeding == the restatement with the same states;
    CropFlipd (one gather) == the two transforms; CenterSpatialCropd == the restatement."""
    from mvtb import spatial as S
    img = P.synthetic_volume(13, (1, 40, 36, 31))
    lab = (P.synthetic_volume(14, (3, 40, 36, 31)) > 0.5).float()
    data = {"image": img.to(cuda_device), "label": lab.to(cuda_device)}
    for seed in range(6):
        crop = S.RandSpatialCropd(["image", "label"], roi_size=[16, 20, 12], random_size=False)
        fl = S.RandFlipd(["image", "label"], prob=0.5, spatial_axis=0)
        crop.set_random_state(seed=100 + seed); fl.set_random_state(seed=200 + seed)
        out = fl(crop(data))
        Rc, Rf = np.random.RandomState(100 + seed), np.random.RandomState(200 + seed)
        sl = M.rand_spatial_crop_slices((40, 36, 31), [16, 20, 12], Rc)
        do = Rf.rand() < 0.5
        for k, src in (("image", img), ("label", lab)):
            want = M.crop_then_flip(src.numpy(), sl, do, 0)
            assert torch.equal(out[k].cpu(), torch.from_numpy(want))
        both = S.CropFlipd(["image", "label"], roi_size=[16, 20, 12], prob=0.5, spatial_axis=0)
        both.crop.set_random_state(seed=100 + seed); both.flipper.set_random_state(seed=200 + seed)
        o2 = both(data)
        assert torch.equal(o2["image"], out["image"]) and torch.equal(o2["label"], out["label"])
    c = S.CenterSpatialCropd(["image", "label"], roi_size=[16, 20, 12])(data)
    sl = M.center_crop_slices((40, 36, 31), [16, 20, 12])
    assert torch.equal(c["image"].cpu(), torch.from_numpy(np.ascontiguousarray(img.numpy()[(slice(None),) + sl])))
    # random_size=True (MONAI's default): size drawn per axis first, then the corner
    rs = S.RandSpatialCropd("image", roi_size=[16, 20, 12], random_size=True)
    rs.set_random_state(seed=5)
    o = rs(data)
    R = np.random.RandomState(5)
    sl = M.rand_spatial_crop_slices((40, 36, 31), [16, 20, 12], R, random_size=True)
    assert torch.equal(o["image"].cpu(), torch.from_numpy(np.ascontiguousarray(img.numpy()[(slice(None),) + sl])))


@pytest.mark.gpu
def test_gpu_batched_gather(cuda_device):
    from mvtb import spatial as S
    x = torch.stack([P.synthetic_volume(20 + b, (2, 20, 18, 37)) for b in range(5)])
    starts = [(0, 0, 0), (4, 2, 5), (7, 0, 1), (1, 3, 0), (8, 6, 5)]
    masks = [0, 1, 6, 7, 4]
    got = S.crop_flip_batch(x.to(cuda_device), starts, (12, 12, 32), masks)
    for b in range(5):
        assert torch.equal(got[b].cpu(), torch.from_numpy(_want(x[b].numpy(), (12, 12, 32), starts[b], masks[b])))
    with pytest.raises(ValueError):
        S.crop_flip_batch(x.to(cuda_device), [(9, 0, 0)] * 5, (12, 12, 32), masks)
