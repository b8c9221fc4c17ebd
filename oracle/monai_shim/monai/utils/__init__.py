from typing import Any, Tuple


def _iterable(v: Any) -> bool:
    try:
        if hasattr(v, "ndim") and v.ndim == 0:
            return False
        iter(v)
        return True
    except TypeError:
        return False


def ensure_tuple(vals: Any) -> Tuple[Any, ...]:
    if isinstance(vals, str) or not _iterable(vals):
        return (vals,)
    return tuple(vals)
