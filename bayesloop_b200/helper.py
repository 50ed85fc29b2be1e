"""Grid builders and nested-list utilities used by the host side of the engine.

`oint`/`cint` define the exact grid coordinates (reference: bayesloop/helper.py:90-120); the nested-list
helpers serve the hyper-parameter plumbing (reference: bayesloop/helper.py:11-62).
"""
import numpy as np

__all__ = ["oint", "cint", "flatten", "recursiveIndex", "assignNestedItem"]


def oint(start, stop, num):
    """`num` equally spaced values strictly inside (start, stop) -- an open interval."""
    return np.linspace(start, stop, int(num) + 2)[1:-1]


def cint(start, stop, num):
    """`num` equally spaced values on [start, stop] -- a closed interval."""
    return np.linspace(start, stop, int(num))


def flatten(nested):
    """Depth-first generator over the leaves of arbitrarily nested lists/tuples."""
    stack = [iter(nested)]
    while stack:
        try:
            item = next(stack[-1])
        except StopIteration:
            stack.pop()
            continue
        if isinstance(item, (list, tuple)):
            stack.append(iter(item))
        else:
            yield item


def recursiveIndex(nested, query):
    """Index path of the first leaf equal to `query` (empty list if absent)."""
    for pos, item in enumerate(nested):
        if isinstance(item, (list, tuple)):
            sub = recursiveIndex(item, query)
            if sub:
                return [pos] + sub
        elif _same(item, query):
            return [pos]
    return []


def _same(a, b):
    try:
        return bool(a == b)
    except Exception:  # element-wise comparisons of arrays
        return False


def assignNestedItem(nested, path, value):
    """In-place assignment `nested[path[0]][path[1]]... = value`."""
    target = nested
    for pos in path[:-1]:
        target = target[pos]
    target[path[-1]] = value


def is_regular(values, tol=1e-10):
    """True if `values` are equally spaced to within `tol` (second differences), cf. core.py:161, :1174."""
    v = np.asarray(values, dtype=float)
    if v.size < 3:
        return True
    return not np.any(np.abs(np.diff(v, 2)) > tol)
