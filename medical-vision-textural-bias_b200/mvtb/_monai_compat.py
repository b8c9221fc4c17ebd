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


# ----------------------------------------------------------------------------- UNet stand-in
# The reference's Gibbs_UNet / Spikes_UNet (stylization_layers.py:119-139, 154-174) put an artifact layer in front of
# monai.networks.nets.UNet.  The network is a CONSUMER of the hot path, not part of it ("use MONAI's /
# a stub"), so when MONAI is importable the real class is used; otherwise this plain-torch residual U-Net with the
# same constructor arguments and the same layer plan (strided residual units down, transposed convolutions + skip
# concatenation up, instance norm + PReLU) keeps the two compositions constructible and runnable.  Weights are NOT
# interchangeable with MONAI checkpoints.
try:  # pragma: no cover
    from monai.networks.nets import UNet  # type: ignore
    HAVE_MONAI_UNET = True
except Exception:  # noqa: BLE001
    HAVE_MONAI_UNET = False
    import torch
    import torch.nn as nn

    def _conv(dim, transposed=False):
        return {(1, False): nn.Conv1d, (2, False): nn.Conv2d, (3, False): nn.Conv3d,
                (1, True): nn.ConvTranspose1d, (2, True): nn.ConvTranspose2d, (3, True): nn.ConvTranspose3d}[(dim, transposed)]

    def _norm(dim):
        return {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[dim]

    class _ResUnit(nn.Module):
        def __init__(self, dim, cin, cout, stride, subunits, last=False):
            super().__init__()
            layers, c, s = [], cin, stride
            for u in range(max(1, subunits)):
                layers.append(_conv(dim)(c, cout, 3, s, 1))
                if not (last and u == max(1, subunits) - 1):
                    layers += [_norm(dim)(cout), nn.PReLU()]
                c, s = cout, 1
            self.body = nn.Sequential(*layers)
            self.skip = nn.Identity() if (cin == cout and stride == 1) else _conv(dim)(cin, cout, 3 if stride != 1 else 1, stride, 1 if stride != 1 else 0)

        def forward(self, x):
            return self.body(x) + self.skip(x)

    class _Up(nn.Module):
        def __init__(self, dim, cin, cout, stride, res_units, last):
            super().__init__()
            self.up = _conv(dim, True)(cin, cout, 3, stride, 1, output_padding=stride - 1)
            self.post = nn.Sequential() if last else nn.Sequential(_norm(dim)(cout), nn.PReLU())
            self.res = _ResUnit(dim, cout, cout, 1, 1, last=last) if res_units > 0 else nn.Identity()

        def forward(self, x):
            return self.res(self.post(self.up(x)))

    class UNet(nn.Module):  # type: ignore[no-redef]
        """Residual U-Net with monai.networks.nets.UNet's (0.5) constructor: UNet(dimensions, in_channels,
        out_channels, channels, strides, num_res_units=0)."""

        def __init__(self, dimensions, in_channels, out_channels, channels, strides, kernel_size=3, up_kernel_size=3,
                     num_res_units=0, **_ignored):
            super().__init__()
            if len(channels) < 2 or len(strides) != len(channels) - 1:
                raise ValueError("the length of `strides` should equal `len(channels) - 1`")
            self.dimensions = dimensions
            c = list(channels)
            self.down = nn.ModuleList()
            cin = in_channels
            for i in range(len(c) - 1):
                self.down.append(_ResUnit(dimensions, cin, c[i], strides[i], max(1, num_res_units)))
                cin = c[i]
            self.bottom = _ResUnit(dimensions, c[-2], c[-1], 1, max(1, num_res_units))
            self.upl = nn.ModuleList()
            for i in range(len(c) - 2, -1, -1):                    # deepest first
                up_in = c[i] + c[i + 1] if i == len(c) - 2 else c[i] + c[i]
                up_out = out_channels if i == 0 else c[i - 1]
                self.upl.append(_Up(dimensions, up_in, up_out, strides[i], num_res_units, last=(i == 0)))

        def forward(self, x):
            skips = []
            for d in self.down:
                x = d(x)
                skips.append(x)
            x = self.bottom(x)
            for u in self.upl:
                x = u(torch.cat([skips.pop(), x], dim=1))
            return x
