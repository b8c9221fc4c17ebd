"""The intensity prologue of the reference's training pipelines on the GPU: drop-ins for the three MONAI 0.5 dictionary
transforms that sit directly in front of the k-space chain
(10_scripts/127_.../stylized_gibbs12p5_spikes15_wrap0p5_sap0p05_FLAIR.py:134-136)

    NormalizeIntensityd(keys, nonzero=True, channel_wise=True) -> RandScaleIntensityd(keys, factors, prob)
        -> RandShiftIntensityd(keys, offsets, prob)

with MONAI's constructor arguments, attributes (`factor`, `_offset`, `_do_transform`) and random draw order, plus
`IntensityPrologued`, the three as one transform: one statistics pass over the data and ONE affine map
y = x != 0 ? a x + b : t, which the band-limited chain can apply while it reads the volume
(functional.kspace_chain_ex(..., pre_abt=...)).  Arithmetic is in libmvtb.so (csrc/intensity.cu); no CPU fallback."""
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import functional as Fn
from ._monai_compat import KeysCollection, MapTransform, RandomizableTransform


def _pair(v: Union[float, Sequence[float]]) -> Tuple[float, float]:
    if isinstance(v, (int, float)):
        return (min(-v, v), max(-v, v))
    if len(v) != 2:
        raise AssertionError("factors / offsets should be a number or pair of numbers.")
    return (min(v), max(v))


class NormalizeIntensityd(MapTransform):
    """MONAI 0.5 NormalizeIntensityd for the configuration the reference uses: subtrahend / divisor computed from the
    data.  `nonzero=False` normalises over all voxels (the map then also moves zeros)."""

    def __init__(self, keys: KeysCollection, subtrahend=None, divisor=None, nonzero: bool = False,
                 channel_wise: bool = False, dtype=np.float32, allow_missing_keys: bool = False) -> None:
        super().__init__(keys, allow_missing_keys)
        if subtrahend is not None or divisor is not None:
            raise NotImplementedError("NormalizeIntensityd: given subtrahend / divisor are not used by the reference and not built")
        self.nonzero, self.channel_wise = nonzero, channel_wise

    def __call__(self, data):
        d = dict(data)
        for key in self.key_iterator(d):
            d[key] = Fn.intensity_prologue(d[key], nonzero=self.nonzero, channel_wise=self.channel_wise)
        return d


class RandScaleIntensityd(RandomizableTransform, MapTransform):
    def __init__(self, keys: KeysCollection, factors, prob: float = 0.1, allow_missing_keys: bool = False) -> None:
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        self.factors = _pair(factors)
        self.factor = None

    def randomize(self, data=None) -> None:
        self.factor = self.R.uniform(low=self.factors[0], high=self.factors[1])
        super().randomize(None)

    def __call__(self, data):
        d = dict(data)
        self.randomize()
        if not self._do_transform:
            return d
        for key in self.key_iterator(d):
            d[key] = Fn.intensity_scale_shift(d[key], scale=1.0 + self.factor, shift=0.0)
        return d


class RandShiftIntensityd(RandomizableTransform, MapTransform):
    def __init__(self, keys: KeysCollection, offsets, prob: float = 0.1, allow_missing_keys: bool = False) -> None:
        MapTransform.__init__(self, keys, allow_missing_keys)
        RandomizableTransform.__init__(self, prob)
        self.offsets = _pair(offsets)
        self._offset = None

    def randomize(self, data=None) -> None:
        self._offset = self.R.uniform(low=self.offsets[0], high=self.offsets[1])
        super().randomize(None)

    def __call__(self, data):
        d = dict(data)
        self.randomize()
        if not self._do_transform:
            return d
        for key in self.key_iterator(d):
            d[key] = Fn.intensity_scale_shift(d[key], scale=1.0, shift=self._offset)
        return d


class IntensityPrologued(MapTransform):
    """NormalizeIntensityd(nonzero=True, channel_wise=True) -> RandScaleIntensityd(factors, prob) ->
    RandShiftIntensityd(offsets, prob) as one transform: the same random draws in the same order (the scale and the
    shift part are `.scale` / `.shift`, each with its own `R`, seedable like the separate transforms), one statistics
    pass and one map over the data."""

    def __init__(self, keys: KeysCollection, factors=0.1, offsets=0.1, prob: float = 0.5, allow_missing_keys: bool = False) -> None:
        super().__init__(keys, allow_missing_keys)
        self.scale = RandScaleIntensityd(keys, factors, prob, allow_missing_keys)
        self.shift = RandShiftIntensityd(keys, offsets, prob, allow_missing_keys)

    def set_random_state(self, seed: Optional[int] = None, state: Optional[np.random.RandomState] = None):
        R = np.random.RandomState(seed) if seed is not None else (state or np.random.RandomState())
        self.scale.set_random_state(seed=int(R.randint(2 ** 32 - 1, dtype=np.uint32)))      # Compose seeds its members one by one
        self.shift.set_random_state(seed=int(R.randint(2 ** 32 - 1, dtype=np.uint32)))
        return self

    def __call__(self, data):
        d = dict(data)
        self.scale.randomize()
        self.shift.randomize()
        s = 1.0 + self.scale.factor if self.scale._do_transform else 1.0
        t = self.shift._offset if self.shift._do_transform else 0.0
        for key in self.key_iterator(d):
            d[key] = Fn.intensity_prologue(d[key], nonzero=True, channel_wise=True, scale=s, shift=t)
        return d
