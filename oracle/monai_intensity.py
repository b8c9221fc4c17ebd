"""CPU restatement of the MONAI 0.5 intensity transforms that precede the k-space chain.  TEST INFRASTRUCTURE ONLY.

The reference's training scripts run, directly in front of the hot path
(10_scripts/127_gibbs_spikes_wraparound_sap_OneChannel/stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:134-136),

    NormalizeIntensityd(keys="image", nonzero=True, channel_wise=True)
    RandScaleIntensityd(keys="image", factors=0.1, prob=0.5)
    RandShiftIntensityd(keys="image", offsets=0.1, prob=0.5)

These are third-party code that is NOT part of /root/reference: MONAI 0.5.dev2113 (the version the reference's
notebooks print; no lock file pins it), monai/transforms/intensity/array.py.  MONAI is not installed in this image and
cannot be fetched, so the published algorithm is restated here in numpy float32, op for op:

  NormalizeIntensity._normalize(img):   slices = img != 0            (nonzero=True)
                                        if not any(slices): return img
                                        img[slices] = img[slices] - mean(img[slices])
                                        d = std(img[slices])  (population std of the shifted values); d == 0 -> 1
                                        img[slices] = img[slices] / d
                      channel_wise=True: applied to every img[c] on its own; output dtype float32
  ScaleIntensity(factor=f):             img * (1 + f)        (float32)
  ShiftIntensity(offset=o):             img + o              (float32)
  RandScaleIntensity.randomize:         self.factor = R.uniform(lo, hi); then the prob gate R.rand() < prob
  RandShiftIntensity.randomize:         self._offset = R.uniform(lo, hi); then the prob gate

Parity is "unpinned" in the sense of the task statement for this one dependency (no MONAI source or golden vector is
available here); the known-answer checks in tests/test_intensity_prologue.py (masked mean 0 / std 1, zeros untouched,
hand-computed 2 x 2 x 2 case) pin the restatement to the documented behaviour.
"""
import numpy as np


def normalize_intensity(img: np.ndarray, nonzero: bool = True, channel_wise: bool = True) -> np.ndarray:
    img = np.array(img, dtype=np.float32, copy=True)

    def _norm(a):
        slices = (a != 0) if nonzero else np.ones(a.shape, dtype=bool)
        if not np.any(slices):
            return a
        a[slices] = a[slices] - np.mean(a[slices])
        d = np.std(a[slices])
        if d == 0.0:
            d = np.float32(1.0)
        a[slices] = a[slices] / d
        return a

    if channel_wise:
        for c in range(img.shape[0]):
            img[c] = _norm(img[c])
        return img
    return _norm(img)


def scale_intensity(img: np.ndarray, factor: float) -> np.ndarray:
    return (np.asarray(img, dtype=np.float32) * np.float32(1 + factor)).astype(np.float32)


def shift_intensity(img: np.ndarray, offset: float) -> np.ndarray:
    return (np.asarray(img, dtype=np.float32) + np.float32(offset)).astype(np.float32)


def draw_scale_shift(R_scale: np.random.RandomState, R_shift: np.random.RandomState, factors=0.1, offsets=0.1,
                     prob_scale=0.5, prob_shift=0.5):
    """(factor or None, offset or None) in the draw order of RandScaleIntensity / RandShiftIntensity."""
    f = R_scale.uniform(low=min(-factors, factors), high=max(-factors, factors))
    do_f = R_scale.rand() < prob_scale
    o = R_shift.uniform(low=min(-offsets, offsets), high=max(-offsets, offsets))
    do_o = R_shift.rand() < prob_shift
    return (f if do_f else None), (o if do_o else None)


def prologue(img: np.ndarray, factor=None, offset=None) -> np.ndarray:
    y = normalize_intensity(img, nonzero=True, channel_wise=True)
    if factor is not None:
        y = scale_intensity(y, factor)
    if offset is not None:
        y = shift_intensity(y, offset)
    return y
