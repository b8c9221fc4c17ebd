"""The spatial part of the reference's training pipelines on the GPU: drop-ins for the MONAI 0.5 dictionary transforms between
the resampled volume and the intensity prologue
(10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:130-133, :153)

    RandSpatialCropd(keys, roi_size, random_center=True, random_size=True) -> RandFlipd(keys, prob, spatial_axis)
    CenterSpatialCropd(keys, roi_size)

with MONAI's constructor arguments and random draw order (one `R.randint(0, N - roi + 1)` per axis that is larger than the
roi, in axis order; one `R.random() < prob` for the flip), and `CropFlipd`, the two as one gather.  The data movement is
`mvtb_crop_flip_f32` in libmvtb.so (csrc/spatial.cu); no CPU fallback.  Every key of a sample shares the drawn window and
flip, as in MONAI (the window is drawn from the first key's shape)."""
import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, functional as Fn
from ._monai_compat import KeysCollection, MapTransform, Randomizable, RandomizableTransform


def _fall_back(roi_size, img_size: Sequence[int]) -> Tuple[int, ...]:
    nd = len(img_size)
    user = tuple(roi_size) if isinstance(roi_size, (list, tuple, np.ndarray)) else (roi_size,) * nd
    if len(user) != nd:
        raise ValueError(f"roi_size has {len(user)} entries for {nd} spatial axes")
    return tuple(int(d) if (u is None or u <= 0) else int(u) for u, d in zip(user, img_size))


def _flip_mask(spatial_axis: Optional[Union[int, Sequence[int]]], nd: int) -> int:
    if spatial_axis is None:
        axes = range(nd)
    elif isinstance(spatial_axis, int):
        axes = (spatial_axis,)
    else:
        axes = tuple(spatial_axis)
    m = 0
    for a in axes:
        a = a + nd if a < 0 else a
        if not 0 <= a < nd:
            raise ValueError(f"spatial_axis {a} outside the {nd} spatial axes")
        m |= 1 << a
    return m


def crop_flip(x, start: Sequence[int], size: Sequence[int], flip_axes_mask: int = 0):
    """out = flip(x[:, start : start + size]) for a (C, H, W, D) (or (C, H, W)) sample; result on x's device, in x's precision."""
    xd, org = Fn.to_device(x)
    nd = xd.dim() - 1
    if nd not in (2, 3):
        raise ValueError(f"crop_flip expects a (C, H, W[, D]) sample, got rank {xd.dim()}")
    shp = tuple(xd.shape[1:]) if nd == 3 else (1,) + tuple(xd.shape[1:])          # 2-D: a leading axis of length 1
    st = tuple(int(v) for v in start) if nd == 3 else (0,) + tuple(int(v) for v in start)
    sz = tuple(int(v) for v in size) if nd == 3 else (1,) + tuple(int(v) for v in size)
    fm = flip_axes_mask if nd == 3 else flip_axes_mask << 1
    out = torch.empty((xd.shape[0],) + (sz if nd == 3 else sz[1:]), dtype=torch.float32, device=xd.device)
    L = _lib.lib()
    i3 = C.c_int32 * 3
    with torch.cuda.device(xd.device):
        rc = L.mvtb_crop_flip_f32(Fn._ptr(xd), Fn._ptr(out), int(xd.shape[0]), i3(*shp), i3(*sz), i3(*st), int(fm), Fn._stream(xd.device))
    _lib.check(L, rc)
    return Fn.back(out, org)


def crop_flip_batch(x: torch.Tensor, starts, size: Sequence[int], flip_masks) -> torch.Tensor:
    """out[b] = flip_b(x[b, :, start_b : start_b + size]) for a CUDA batch (B, C, H, W, D) in one launch; starts: (B, 3) ints,
    flip_masks: (B,) ints (bit a = spatial axis a), as tensors or sequences."""
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 5:
        raise ValueError("crop_flip_batch expects a contiguous float32 CUDA tensor (B, C, H, W, D)")
    B_, C_ = int(x.shape[0]), int(x.shape[1])
    st = torch.as_tensor(starts, dtype=torch.int32).reshape(B_, 3)
    fm = torch.as_tensor(flip_masks, dtype=torch.int32).reshape(B_)
    sz = tuple(int(v) for v in size)
    if bool((st < 0).any()) or any(int(st[:, a].max()) + sz[a] > int(x.shape[2 + a]) for a in range(3)):
        raise ValueError("crop_flip_batch: a window does not fit in the volume")
    st, fm = st.to(x.device), fm.to(x.device)
    out = torch.empty((B_, C_) + sz, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    i3 = C.c_int32 * 3
    with torch.cuda.device(x.device):
        rc = L.mvtb_crop_flip_batch_f32(Fn._ptr(x), Fn._ptr(out), B_, C_, i3(*[int(v) for v in x.shape[2:]]), i3(*sz), Fn._ptr(st), Fn._ptr(fm),
                                        Fn._stream(x.device))
    _lib.check(L, rc)
    return out


def _center_start(img_size: Sequence[int], roi: Sequence[int]):
    """CenterSpatialCrop -> SpatialCrop(roi_center=[i // 2], roi_size): start = max(center - roi // 2, 0), clipped at the end"""
    start = [max(n // 2 - r // 2, 0) for n, r in zip(img_size, roi)]
    size = [min(s + r, n) - s for s, r, n in zip(start, roi, img_size)]
    return start, size


class CenterSpatialCropd(MapTransform):
    def __init__(self, keys: KeysCollection, roi_size, allow_missing_keys: bool = False) -> None:
        super().__init__(keys, allow_missing_keys)
        self.roi_size = roi_size

    def __call__(self, data):
        d = dict(data)
        for key in self.key_iterator(d):
            img_size = tuple(d[key].shape[1:])
            start, size = _center_start(img_size, _fall_back(self.roi_size, img_size))
            d[key] = crop_flip(d[key], start, size, 0)
        return d


class RandSpatialCropd(Randomizable, MapTransform):
    def __init__(self, keys: KeysCollection, roi_size, random_center: bool = True, random_size: bool = True,
                 allow_missing_keys: bool = False) -> None:
        MapTransform.__init__(self, keys, allow_missing_keys)
        self.roi_size, self.random_center, self.random_size = roi_size, random_center, random_size
        self._start: Optional[Sequence[int]] = None
        self._size: Optional[Sequence[int]] = None

    def randomize(self, img_size: Sequence[int]) -> None:
        self._size = _fall_back(self.roi_size, img_size)
        if self.random_size:
            self._size = tuple(self.R.randint(low=self._size[i], high=img_size[i] + 1) for i in range(len(img_size)))
        if self.random_center:
            valid = tuple(min(ms, ps or ms) for ms, ps in zip(img_size, self._size))
            self._start = tuple(self.R.randint(low=0, high=ms - ps + 1) if ms > ps else 0 for ms, ps in zip(img_size, valid))
            self._size = valid
        else:
            self._start, self._size = _center_start(img_size, self._size)

    def __call__(self, data):
        d = dict(data)
        self.randomize(tuple(d[self.keys[0]].shape[1:]))             # the first key's shape, as in MONAI
        for key in self.key_iterator(d):
            d[key] = crop_flip(d[key], self._start, self._size, 0)
        return d


class RandFlipd(RandomizableTransform, MapTransform):
    def __init__(self, keys: KeysCollection, prob: float = 0.1, spatial_axis: Optional[Union[Sequence[int], int]] = None,
                 allow_missing_keys: bool = False) -> None:
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        self.spatial_axis = spatial_axis

    def __call__(self, data):
        self.randomize(None)
        d = dict(data)
        if not self._do_transform:
            return d
        for key in self.key_iterator(d):
            nd = d[key].dim() - 1 if isinstance(d[key], torch.Tensor) else np.ndim(d[key]) - 1
            d[key] = crop_flip(d[key], (0,) * nd, tuple(d[key].shape[1:]), _flip_mask(self.spatial_axis, nd))
        return d


class CropFlipd(Randomizable, MapTransform):
    """RandSpatialCropd(random_size=False) followed by RandFlipd as ONE pass over the crop: the same draws in the same order
    from two random states (set_random_state seeds both the way Compose seeds two consecutive transforms)."""

    def __init__(self, keys: KeysCollection, roi_size, prob: float = 0.5, spatial_axis=0, allow_missing_keys: bool = False) -> None:
        MapTransform.__init__(self, keys, allow_missing_keys)
        self.crop = RandSpatialCropd(keys, roi_size, random_center=True, random_size=False, allow_missing_keys=allow_missing_keys)
        self.flipper = RandFlipd(keys, prob, spatial_axis, allow_missing_keys=allow_missing_keys)

    def set_random_state(self, seed: Optional[int] = None, state: Optional[np.random.RandomState] = None):
        Randomizable.set_random_state(self, seed, state)
        MAX_SEED = np.iinfo(np.uint32).max + 1
        self.crop.set_random_state(seed=int(self.R.randint(MAX_SEED, dtype="uint32")))
        self.flipper.set_random_state(seed=int(self.R.randint(MAX_SEED, dtype="uint32")))
        return self

    def randomize(self, data=None) -> None:
        pass

    def __call__(self, data):
        d = dict(data)
        self.crop.randomize(tuple(d[self.keys[0]].shape[1:]))
        self.flipper.randomize(None)
        for key in self.key_iterator(d):
            nd = len(self.crop._size)
            fm = _flip_mask(self.flipper.spatial_axis, nd) if self.flipper._do_transform else 0
            d[key] = crop_flip(d[key], self.crop._start, self.crop._size, fm)
        return d
