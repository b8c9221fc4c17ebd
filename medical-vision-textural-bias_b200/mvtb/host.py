"""Host-side logic of the drop-in transforms: everything that is not voxel arithmetic.

Bit-exact masks are obtained by turning each of the reference's floating-point mask
predicates into an integer threshold on an exactly representable squared distance
(SURVEY.md A.1); the kernels then compare integers.  No CUDA is needed to import this
module, so `-m "not gpu"` tests cover it.

F = source_code/filters_and_operators.py, S = source_code/stylization_layers.py of the reference.
"""
import ctypes
from functools import lru_cache
from math import floor
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_BIG = 1 << 62


def _largest_true(pred, hi: int) -> int:
    """Largest q in [0, hi] with pred(q) for a predicate that is True then False; -1 if none."""
    if not pred(0):
        return -1
    if pred(hi):
        return hi
    lo = 0                      # pred(lo) True, pred(hi) False
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if pred(mid):
            lo = mid
        else:
            hi = mid
    return lo


def disk_threshold(r: float, shape_tail: Sequence[int]) -> int:
    """keep <=> sum (i - floor(N/2))^2 < r**2, evaluated by torch in float32 (F:184-187).

    Returns thr such that keep <=> sum <= thr (thr = -1 keeps nothing).  Memoised: a transform asks for the same
    (r, shape) on every call, and the bisection costs ~15 us of the ~100 us a one-volume call spends on the host."""
    if isinstance(r, (int, float)):
        return _disk_threshold_cached(float(r), tuple(int(n) for n in shape_tail))
    return _disk_threshold(r, shape_tail)


@lru_cache(maxsize=256)
def _disk_threshold_cached(r: float, shape_tail: tuple) -> int:
    return _disk_threshold(r, shape_tail)


def _disk_threshold(r, shape_tail: Sequence[int]) -> int:
    r2 = np.float32(float(r) ** 2) if not isinstance(r, complex) else np.float32(np.nan)
    smax = int(sum(max(floor(n / 2), n - 1 - floor(n / 2)) ** 2 for n in shape_tail))
    return _largest_true(lambda s: bool(np.float32(s) < r2), smax)


def _centred_qmax(shape: Sequence[int]) -> int:
    return int(sum((n - 1) ** 2 for n in shape))


def gibbs_threshold(alpha: float, shape: Sequence[int]) -> int:
    """GibbsNoise mask (F:686-698): keep <=> sqrt(sum (i-(N-1)/2)^2) <= r in numpy float64,
    r = (1-alpha) * max(shape) * sqrt(2) / 2.  With q = sum (2i-(N-1))^2 the distance is sqrt(q/4)."""
    shape = tuple(int(s) for s in shape)
    r = (1 - alpha) * np.max(shape) * np.sqrt(2) / 2.0
    return _largest_true(lambda q: bool(np.sqrt(np.float64(q) / 4.0) <= r), _centred_qmax(shape))


def layer_threshold(alpha, shape: Sequence[int]) -> int:
    """GibbsNoiseLayer mask (S:99-109), all float32: keep <=> not (dist / (alpha * dist.max()) > 1)."""
    shape = tuple(int(s) for s in shape)
    a = np.float32(float(alpha))
    qmax = _centred_qmax(shape)
    with np.errstate(divide="ignore", invalid="ignore"):
        dmax = np.sqrt(np.float32(qmax) / np.float32(4.0), dtype=np.float32)
        an = np.float32(a * dmax)

        def keep(q):
            dist = np.sqrt(np.float32(q) / np.float32(4.0), dtype=np.float32)
            return not bool(np.float32(dist / an) > np.float32(1.0))

        return _largest_true(keep, qmax)


@lru_cache(maxsize=32)
def ellipsoid_shell(shape3: Tuple[int, int, int], a: float, b: float, c: float) -> np.ndarray:
    """Row-major (n,3) list of fftshift-ed indices on the shell .95 < sum((i-c)^2/a^2) < 1.05.

    Same float32 expression tree as the reference (F:307-315: int64 squares divided by Python
    floats, summed left to right); it depends only on (shape, a, b, c), so it is computed once
    and cached instead of once per call."""
    h, w, d = (int(s) for s in shape3)
    ih = torch.arange(0, h) - floor(h / 2)
    iw = torch.arange(0, w) - floor(w / 2)
    idd = torch.arange(0, d) - floor(d / 2)
    t = (ih[:, None, None] ** 2) / a ** 2 + (iw[None, :, None] ** 2) / b ** 2 + (idd[None, None, :] ** 2) / c ** 2
    sel = torch.logical_and(t > .95, t < 1.05)
    return sel.nonzero().numpy()


def exp_f32(log_intensity: float) -> float:
    """exp() of a log-intensity that the reference stores in a float32 tensor (F:387-389, F:942)."""
    return float(torch.exp(torch.tensor(float(log_intensity), dtype=torch.float32)))


def make_desc(*, mask_kind: int = _lib.MASK_NONE, mask_ndim: int = 0, mask_thresh: int = 0, inside_off: bool = False,
              spikes: Iterable[Tuple[Sequence[int], float]] = (), wrap_alpha: Optional[float] = None,
              wrap_naxes: int = 3, mask_u: int = 0, mask_p: float = 0.0) -> _lib.ChainDesc:
    """Build one mvtb_chain_desc.  spikes: (fftshift-ed index per FFT axis outermost first, amplitude);
    later entries at the same location replace earlier ones (the reference overwrites, F:937-938)."""
    d = _lib.ChainDesc()
    d.mask_kind = int(mask_kind)
    d.mask_ndim = int(mask_ndim)
    d.mask_thresh = int(max(min(mask_thresh, _BIG), -1))
    d.inside_off = 1 if inside_off else 0
    uniq = {}
    for idx, amp in spikes:
        uniq[tuple(int(i) for i in idx)] = float(amp)
    if len(uniq) > _lib.MAX_SPIKES:
        raise ValueError(f"at most {_lib.MAX_SPIKES} spike locations per volume are supported, got {len(uniq)}")
    d.n_spikes = len(uniq)
    for s, (idx, amp) in enumerate(uniq.items()):
        for j, i in enumerate(idx):
            d.spikes[s].idx[j] = i
        d.spikes[s].amplitude = amp
    d.mask_u = int(mask_u) or None           # MASK_UNIFORM: device address of the volume's uniform field (the caller keeps it alive)
    d.mask_p = float(mask_p)
    if wrap_alpha is None:
        d.wrap_alpha, d.wrap_naxes = 1.0, 0
    else:
        d.wrap_alpha, d.wrap_naxes = float(wrap_alpha), int(wrap_naxes)
    return d


def desc_array(descs: Sequence[_lib.ChainDesc]):
    """The contiguous ChainDesc array the C ABI takes; an array made earlier is passed through as it is."""
    if isinstance(descs, ctypes.Array):
        return descs
    arr = (_lib.ChainDesc * len(descs))()
    for i, d in enumerate(descs):
        arr[i] = d
    return arr
