"""Transform protocol stand-ins (MONAI 0.5 semantics)."""
from typing import Any, Hashable, Optional

import numpy as np

from monai.utils import ensure_tuple


class Transform:
    def __call__(self, data: Any):
        raise NotImplementedError


class Randomizable:
    # class-level stream shared by every instance that was never seeded
    R: np.random.RandomState = np.random.RandomState()

    def set_random_state(self, seed: Optional[int] = None, state: Optional[np.random.RandomState] = None):
        if seed is not None:
            self.R = np.random.RandomState(int(seed) % (2 ** 32))
            return self
        if state is not None:
            if not isinstance(state, np.random.RandomState):
                raise TypeError("state must be a numpy RandomState")
            self.R = state
            return self
        self.R = np.random.RandomState()
        return self

    def randomize(self, data: Any) -> None:
        raise NotImplementedError


class RandomizableTransform(Randomizable, Transform):
    def __init__(self, prob: float = 1.0, do_transform: bool = True):
        self._do_transform = do_transform
        self.prob = min(max(prob, 0.0), 1.0)

    def randomize(self, data: Any) -> None:
        self._do_transform = self.R.rand() < self.prob


class MapTransform(Transform):
    def __init__(self, keys, allow_missing_keys: bool = False) -> None:
        self.keys = ensure_tuple(keys)
        self.allow_missing_keys = allow_missing_keys
        if not self.keys:
            raise ValueError("keys must be non empty.")
        for key in self.keys:
            if not isinstance(key, Hashable):
                raise TypeError(f"keys must be one of (Hashable, Iterable[Hashable]) but is {type(keys).__name__}.")

    def key_iterator(self, data, *extra_iterables):
        ex_iters = extra_iterables if extra_iterables else [[None] * len(self.keys)]
        for key, *_ex in zip(self.keys, *ex_iters):
            if key in data:
                yield (key,) + tuple(_ex) if extra_iterables else key
            elif not self.allow_missing_keys:
                raise KeyError(f"Key was missing ({key}) and allow_missing_keys==False")
