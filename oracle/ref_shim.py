"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (christophmark/bayesloop, mounted read-only at
/root/reference) inside this container so that golden vectors can be generated from it.
The GPU box has no /root/reference: nothing under tests -m gpu / smoke() / bench.py may
import this module at run time; they read the committed fixtures in tests/golden/.

Three in-memory shims are needed on this image:
  * stub `matplotlib`, `matplotlib.pyplot`, `matplotlib.colors` (imported by bayesloop/core.py:13, helper.py:8)
  * stub `pyparsing` (imported by bayesloop/parser.py:9)
  * `numpy.math` (Poisson.pdf calls np.math.factorial, bayesloop/observationModels.py:502; removed in NumPy 2).
    The stand-in factorial also accepts integral floats, as math.factorial did on the Python versions the
    reference was written for (<3.10) -- otherwise any float-typed count series (e.g. one holding NaN for
    missing data) raises TypeError on Python 3.12 before the algorithm even starts.
"""
import math
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("BAYESLOOP_REFERENCE", "/root/reference")


class _Anything:
    """Attribute sink: every attribute/call/subscript yields another sink."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __add__(self, o):
        return _Anything()

    __radd__ = __or__ = __ror__ = __xor__ = __lshift__ = __rshift__ = __mul__ = __rmul__ = __add__
    __sub__ = __and__ = __invert__ = __neg__ = __pow__ = __add__


def _stub_module(name):
    m = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):  # keep inspect / importlib machinery honest (torch walks sys.modules)
            raise AttributeError(attr)
        return _Anything

    m.__getattr__ = _getattr  # type: ignore[attr-defined]
    m.__path__ = []  # behave like a package
    return m


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bayesloop"))


def import_reference():
    """Return the reference `bayesloop` module (raises RuntimeError if not mounted)."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    import numpy as np

    if not hasattr(np, "math"):
        def _factorial(x):
            if isinstance(x, (float, np.floating)):
                if x != int(x):
                    raise ValueError("factorial() only accepts integral values")
                x = int(x)
            return math.factorial(int(x))

        np.math = types.SimpleNamespace(factorial=_factorial)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "pyparsing"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _stub_module(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import bayesloop  # noqa: E402

    return bayesloop
