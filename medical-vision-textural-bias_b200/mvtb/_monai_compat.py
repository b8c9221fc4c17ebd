"""Transform protocol bases.  Real MONAI is used when importable; otherwise these minimal
equivalents of the MONAI 0.5 classes the reference builds on (filters_and_operators.py:11-13)
keep the drop-in modules importable.  Only the behaviour the hot path relies on is provided:
`R` random streams, `set_random_state`, the `prob` gate and key iteration.
"""
from typing import Any, Collection, Hashable, Optional, Tuple, Union

import numpy as np

try:  # pragma: no cover - MONAI is not installed in the build image
    from monai.config import KeysCollection
    from monai.transforms import MapTransform, Randomizable, RandomizableTransform, Transform
    from monai.utils import ensure_tuple
    HAVE_MONAI = True
except Exception:  # noqa: BLE001
    HAVE_MONAI = False
    KeysCollection = Union[Collection[Hashable], Hashable]

    def ensure_tuple(vals: Any) -> Tuple[Any, ...]:
        if isinstance(vals, str):
            return (vals,)
        if getattr(vals, "ndim", None) == 0:
            return (vals,)
        try:
            return tuple(vals)
        except TypeError:
            return (vals,)

    class Transform:
        def __call__(self, data: Any):
            raise NotImplementedError(f"{type(self).__name__} must implement __call__")

    class Randomizable:
        R: np.random.RandomState = np.random.RandomState()   # shared until an instance is seeded

        def set_random_state(self, seed: Optional[int] = None, state: Optional[np.random.RandomState] = None):
            if seed is not None:
                self.R = np.random.RandomState(int(seed) % (1 << 32))
            elif state is not None:
                if not isinstance(state, np.random.RandomState):
                    raise TypeError(f"state must be None or a np.random.RandomState but is {type(state).__name__}.")
                self.R = state
            else:
                self.R = np.random.RandomState()
            return self

        def randomize(self, data: Any) -> None:
            raise NotImplementedError(f"{type(self).__name__} must implement randomize")

    class RandomizableTransform(Randomizable, Transform):
        def __init__(self, prob: float = 1.0, do_transform: bool = True):
            self._do_transform = do_transform
            self.prob = min(max(prob, 0.0), 1.0)

        def randomize(self, data: Any) -> None:
            self._do_transform = self.R.rand() < self.prob

    class MapTransform(Transform):
        def __init__(self, keys, allow_missing_keys: bool = False) -> None:
            self.keys: Tuple[Hashable, ...] = ensure_tuple(keys)
            self.allow_missing_keys = allow_missing_keys
            if not self.keys:
                raise ValueError("keys must be non empty.")
            for k in self.keys:
                if not isinstance(k, Hashable):
                    raise TypeError(f"keys must be one of (Hashable, Iterable[Hashable]) but is {type(keys).__name__}.")

        def key_iterator(self, data, *extra_iterables):
            extras = extra_iterables if extra_iterables else [[None] * len(self.keys)]
            for key, *rest in zip(self.keys, *extras):
                if key in data:
                    yield ((key,) + tuple(rest)) if extra_iterables else key
                elif not self.allow_missing_keys:
                    raise KeyError(f"Key was missing ({key}) and allow_missing_keys==False")
