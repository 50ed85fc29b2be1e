"""Shared comparison of a fitted product study against a golden fixture of the reference."""
import contextlib
import io

import numpy as np

import cases

# keys that hold uninitialised memory in the reference for aborted fits (core.py:360, :399-400, :1356)
GARBAGE_WHEN_DEAD = ('localEvidence', 'posteriorSequence', 'posteriorMeanValues')


def run_case(name, bl):
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), np.errstate(all='ignore'):
        S = cases.CASES[name](bl)
    return S, cases.extract(S)


def compare(name, got, want, rtol, atol_post=0.0):
    """Relative comparison key by key; posterior grids are compared relative to the per-row maximum for cells
    below `floor` (cells that are ~1e-300 of the mode carry no information)."""
    dead = not np.all(np.isfinite(np.atleast_1d(want.get('logEvidenceList', want['logEvidence']))))
    for key, ref in want.items():
        if dead and key in GARBAGE_WHEN_DEAD:
            continue
        assert key in got, '{}: missing result {}'.format(name, key)
        val = np.asarray(got[key], dtype=float)
        ref = np.asarray(ref, dtype=float)
        assert val.shape == ref.shape, '{}: {} shape {} != {}'.format(name, key, val.shape, ref.shape)
        both_nan = np.isnan(val) & np.isnan(ref)
        same_inf = np.isinf(ref) & (val == ref)
        ok = both_nan | same_inf
        scale = np.abs(ref)
        if key in ('posteriorSequence', 'marginalizedPosterior') or key.startswith('parameterPosterior'):
            # tolerance relative to the value, with an absolute floor tied to the row maximum
            rowmax = np.max(np.abs(ref).reshape(ref.shape[0], -1), axis=1).reshape([-1] + [1] * (ref.ndim - 1)) \
                if ref.ndim > 1 else np.max(np.abs(ref))
            tol = rtol * scale + atol_post * rowmax
        else:
            finite = np.abs(ref[np.isfinite(ref)])
            tol = rtol * scale + rtol * (finite.max() if finite.size else 0.0) + 1e-300
        with np.errstate(invalid='ignore'):
            close = np.abs(val - ref) <= tol
        bad = ~(ok | close)
        if bad.any():
            idx = np.unravel_index(np.argmax(bad), bad.shape) if bad.ndim else ()
            raise AssertionError('{}: {} differs at {}: got {!r}, want {!r} (rtol {})'
                                 .format(name, key, idx, val[idx] if bad.ndim else val, ref[idx] if bad.ndim else ref,
                                         rtol))
