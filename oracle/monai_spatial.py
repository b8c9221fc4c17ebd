"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU legs): numpy restatement of the MONAI 0.5 spatial transforms
that sit between the resampled volume and the intensity prologue in the reference's training scripts
(10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:130-133, :153):

    RandSpatialCropd(keys, roi_size=[128, 128, 64], random_size=False) -> RandFlipd(keys, prob=0.5, spatial_axis=0)
    CenterSpatialCropd(keys, roi_size=[128, 128, 64])                     (validation)

MONAI (0.5.dev2113, the version the scripts print) is a third-party dependency that is absent from /root/reference and
from this image, so these functions restate its published algorithm (monai/transforms/croppad/array.py: SpatialCrop,
CenterSpatialCrop, RandSpatialCrop; monai/transforms/utils.py: get_valid_patch_size, get_random_patch;
monai/transforms/spatial/array.py: Flip, RandFlip; monai/utils/misc.py: fall_back_tuple) and are pinned by known-answer
tests only (tests/test_spatial.py): PARITY UNPINNED against MONAI itself.  Data movement is exact by construction."""
from typing import Optional, Sequence, Tuple, Union

import numpy as np


def fall_back_tuple(user_provided, default: Sequence[int]) -> Tuple[int, ...]:
    """non-positive / None entries of roi_size fall back to the image size"""
    nd = len(default)
    user = tuple(user_provided) if isinstance(user_provided, (list, tuple, np.ndarray)) else (user_provided,) * nd
    if len(user) != nd:
        raise ValueError(f"roi_size has {len(user)} entries for {nd} spatial axes")
    return tuple(int(d) if (u is None or u <= 0) else int(u) for u, d in zip(user, default))


def get_valid_patch_size(image_size: Sequence[int], patch_size) -> Tuple[int, ...]:
    patch = fall_back_tuple(patch_size, image_size) if not isinstance(patch_size, tuple) else patch_size
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch))


def get_random_patch(dims: Sequence[int], patch_size: Sequence[int], R: np.random.RandomState) -> Tuple[slice, ...]:
    """one R.randint(0, ms - ps + 1) per axis whose image is larger than the patch, in axis order"""
    min_corner = tuple(R.randint(low=0, high=ms - ps + 1) if ms > ps else 0 for ms, ps in zip(dims, patch_size))
    return tuple(slice(mc, mc + ps) for mc, ps in zip(min_corner, patch_size))


def center_crop_slices(img_size: Sequence[int], roi_size) -> Tuple[slice, ...]:
    """CenterSpatialCrop -> SpatialCrop(roi_center=[i // 2], roi_size): start = max(center - roi // 2, 0), end = start + roi"""
    roi = np.asarray(fall_back_tuple(roi_size, img_size), dtype=np.int16)
    center = np.asarray([i // 2 for i in img_size], dtype=np.int16)
    start = np.maximum(center - np.floor_divide(roi, 2), 0)
    end = np.maximum(start + roi, start)
    return tuple(slice(int(s), int(min(e, n))) for s, e, n in zip(start, end, img_size))


def rand_spatial_crop_slices(img_size: Sequence[int], roi_size, R: np.random.RandomState, random_center: bool = True,
                             random_size: bool = False) -> Tuple[slice, ...]:
    size = fall_back_tuple(roi_size, img_size)
    if random_size:
        size = tuple(R.randint(low=size[i], high=img_size[i] + 1) for i in range(len(img_size)))
    if random_center:
        valid = get_valid_patch_size(img_size, size)
        return get_random_patch(img_size, valid, R)
    return center_crop_slices(img_size, size)


def flip(img: np.ndarray, spatial_axis: Optional[Union[int, Sequence[int]]]) -> np.ndarray:
    """Flip.__call__: np.flip per channel over spatial_axis (None = every spatial axis)"""
    return np.stack([np.flip(ch, spatial_axis) for ch in img]).astype(img.dtype)


def crop_then_flip(img: np.ndarray, slices: Sequence[slice], do_flip: bool, spatial_axis) -> np.ndarray:
    out = img[(slice(None),) + tuple(slices)]
    return flip(out, spatial_axis) if do_flip else np.ascontiguousarray(out)
