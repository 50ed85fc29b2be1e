"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

NumPy restatement ("port") of the reference's grid forward-backward path, written the way the reference runs it:
one Python loop over time steps, a handful of NumPy passes over the grid per step, SciPy's compiled
`gaussian_filter1d` for the GaussianRandomWalk (the reference calls exactly that: transitionModels.py:111), and
the log-space `np.logaddexp` model averaging of HyperStudy.fit (core.py:1358-1366, :1372-1385).

Used for (1) bench.py's `cpu_baseline` / `--impl reference` legs (timed on the GPU box's host cores; /root/reference
does not exist there), (2) an independent cross-check of oracle/blgrid_oracle.c (linear-space averaging vs the
reference's log-space form) and of the restated reflect filter against SciPy.

Parity status: PINNED -- tests/test_oracle.py runs it against the golden fixtures in tests/golden/ that were
generated from the unmodified reference.

It consumes the same flat "transition program" as the C ABI (include/blgrid.h) so that one lowering feeds the
CUDA library, the C oracle and this port.
"""
import math

import numpy as np

try:  # SciPy 1.18.1 on this image (reference pin: scipy>=0.17.1, setup.py:14)
    from scipy.ndimage import gaussian_filter1d as _scipy_gaussian_filter1d
except Exception:  # pragma: no cover
    _scipy_gaussian_filter1d = None

OM_POISSON, OM_GAUSSIAN, OM_SCALED_AR1, OM_AR1, OM_WHITE_NOISE, OM_GAUSSIAN_MEAN, OM_LAPLACE, OM_BERNOULLI = range(1, 9)
OM_TABLE = 100
OP_GRW, OP_REGIME, OP_RESET, OP_NOTEQUAL = 1, 2, 3, 4


def gaussian_filter1d_reflect(x, sigma, axis, radius=None):
    """Restatement of scipy.ndimage.gaussian_filter1d(x, sigma, axis) with its defaults (order 0, mode='reflect',
    truncate=4.0): kernel exp(-0.5/sigma^2 * j^2), j = -R..R, R = int(4*sigma + 0.5), normalised to sum 1
    (_filters.py:656-666, :747); boundary = half-sample symmetric extension with period 2n (NI_EXTEND_REFLECT)."""
    R = int(4.0 * float(sigma) + 0.5) if radius is None else int(radius)
    j = np.arange(-R, R + 1)
    w = np.exp(-0.5 / (sigma * sigma) * j ** 2)
    w = w / w.sum()
    n = x.shape[axis]
    idx = (np.arange(n)[:, None] + j[None, :]) % (2 * n)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    xm = np.moveaxis(x, axis, -1)
    out = np.tensordot(xm[..., idx], w, axes=([-1], [0]))
    return np.moveaxis(out, -1, axis)


def _filter(x, sigma, axis, use_scipy):
    if use_scipy and _scipy_gaussian_filter1d is not None:
        return _scipy_gaussian_filter1d(x, sigma, axis=axis)
    return gaussian_filter1d_reflect(x, sigma, axis)


def pdf(om, grid, d):
    """The reference's pdf expressions (observationModels.py) on the meshgrid `grid` for one column segment d."""
    if om == OM_POISSON:
        return (grid[0] ** d[0]) * np.exp(-grid[0]) / math.factorial(int(d[0]))
    if om == OM_GAUSSIAN:
        return np.exp(-((d[0] - grid[0]) ** 2.) / (2. * grid[1] ** 2.) - .5 * np.log(2. * np.pi * grid[1] ** 2.))
    if om == OM_SCALED_AR1:
        ss = grid[1] * np.sqrt(1 - grid[0] ** 2.)
        return np.exp(-((d[1] - grid[0] * d[0]) ** 2.) / (2. * ss ** 2.) - .5 * np.log(2. * np.pi * ss ** 2.))
    if om == OM_AR1:
        return np.exp(-((d[1] - grid[0] * d[0]) ** 2.) / (2. * grid[1] ** 2.) - .5 * np.log(2. * np.pi * grid[1] ** 2.))
    if om == OM_WHITE_NOISE:
        return np.exp(-(d[0] ** 2.) / (2. * grid[0] ** 2.) - .5 * np.log(2. * np.pi * grid[0] ** 2.))
    if om == OM_LAPLACE:
        return np.exp(-np.abs(d[0] - grid[0]) / grid[1]) / (2. * grid[1])
    if om == OM_BERNOULLI:
        p = np.where((grid[0] > 1.) | (grid[0] < 0.), 0., grid[0])
        return p if d[0] else 1. - p
    raise ValueError(om)


def processed_pdf(om, grid, segment, multiply=True):
    """observationModels.py:35-56"""
    segment = np.asarray(segment, dtype=float)
    if om == OM_GAUSSIAN_MEAN:
        if np.isnan(segment).any():
            return np.ones_like(grid[0])
        m, s = segment[0, 0], segment[0, 1]
        return np.exp(-((m - grid[0]) ** 2.) / (2. * s ** 2.) - .5 * np.log(2. * np.pi * s ** 2.))
    if segment.ndim == 2 and multiply:
        return np.prod(np.array([processed_pdf(om, grid, col) for col in segment.T]), axis=0)
    if np.isnan(segment).any():
        return np.ones_like(grid[0])
    return pdf(om, grid, segment)


class Problem:
    def __init__(self, coords, lattice, om, seg_len, data, prior, reset_base=None, lik_table=None):
        self.coords = [np.asarray(c, dtype=float) for c in coords]
        self.shape = tuple(len(c) for c in self.coords)
        self.grid = list(np.meshgrid(*self.coords, indexing='ij'))
        self.lattice = [float(x) for x in lattice]
        self.lc = float(np.prod(self.lattice))
        self.om, self.seg = om, seg_len
        raw = np.asarray(data, dtype=float)
        self.segments = np.array([raw[i:i + seg_len] for i in range(raw.shape[0] - (seg_len - 1))])
        self.T = len(self.segments)
        self.prior = np.asarray(prior, dtype=float).reshape(self.shape)
        self.reset_base = None if reset_base is None else np.asarray(reset_base, dtype=float).reshape(self.shape)
        self.lik_table = lik_table

    def likelihood(self, i):
        if self.om == OM_TABLE:
            return np.asarray(self.lik_table[i], dtype=float).reshape(self.shape)
        return processed_pdf(self.om, self.grid, self.segments[i])


def apply_ops(pb, ops, b, idx, backward, x, use_scipy=True):
    """Transition program of one step for combo b (see include/blgrid.h)."""
    for op in ops:
        f_lo, f_hi, b_lo, b_hi = op['window'][b]
        lo, hi = (b_lo, b_hi) if backward else (f_lo, f_hi)
        if idx < lo or idx >= hi:
            continue
        par = op['param'][b]
        kind = op['kind']
        if kind == OP_GRW:  # transitionModels.py:107-113
            x = _filter(x, par, op['axis'], use_scipy) if par > 0. else x.copy()
        elif kind == OP_REGIME:  # :405-410
            x = x.copy()
            x[x < par] = par
            x /= np.sum(x)
        elif kind == OP_RESET:  # :300-312
            x = pb.reset_base * par
        elif kind == OP_NOTEQUAL:  # :461-469
            x = np.amax(x) - x
            x /= np.sum(x)
            x[x < par] = par
            x /= np.sum(x)
    return x


def fit_combo(pb, ops, b, forward_only=False, evidence_only=False, use_scipy=True):
    """Study.fit for one combination (core.py:330-486).  Returns (logE, localEvidence, posteriorSequence|None)."""
    T = pb.T
    post = None if evidence_only else np.empty((T,) + pb.shape)
    local = np.empty(T)
    logE = 0.
    alpha = pb.prior.copy()
    for i in range(T):
        lik = pb.likelihood(i)
        alpha = alpha * lik
        norm = np.sum(alpha)
        if not norm > 0.:
            return -np.inf, local, post
        alpha /= norm
        logE += np.log(norm)
        local[i] = norm * pb.lc
        if post is not None:
            post[i] = alpha
        alpha = apply_ops(pb, ops, b, i, False, alpha, use_scipy)
    logE += np.log(pb.lc)
    if not (forward_only or evidence_only):
        beta = np.ones(pb.shape)
        beta /= np.sum(beta)
        for i in range(T - 1, -1, -1):
            post[i] *= beta
            norm = np.sum(post[i])
            if not norm > 0.:
                return -np.inf, local, post
            post[i] /= np.sum(post[i])
            lik = pb.likelihood(i)
            with np.errstate(invalid='ignore', divide='ignore'):
                local[i] = 1. / (np.sum(post[i] / lik) * pb.lc)
            beta = apply_ops(pb, ops, b, i, True, beta * lik, use_scipy)
            beta /= np.sum(beta)
    return logE, local, post


def hyper_fit(pb, ops, hyper_prior, forward_only=False, evidence_only=False, use_scipy=True, rows=None):
    """HyperStudy.fit's combo loop and log-space averaging (core.py:1349-1419).  Returns a dict of host arrays."""
    rows = range(len(hyper_prior)) if rows is None else rows
    avg = None if evidence_only else np.zeros((pb.T,) + pb.shape) - np.inf
    logEs, locals_ = [], []
    for b in rows:
        logE, local, post = fit_combo(pb, ops, b, forward_only, evidence_only, use_scipy)
        logEs.append(logE)
        locals_.append(local)
        if (not evidence_only) and np.isfinite(logE):
            post[post < 10. ** -300] = 10. ** -300
            with np.errstate(divide='ignore'):
                avg = np.logaddexp(avg, np.log(post) + logE + np.log(hyper_prior[b]))
    out = dict(logEvidenceList=np.array(logEs), localEvidenceList=np.array(locals_), logAverage=avg)
    return out


def finish_average(pb, log_avg):
    """core.py:1372-1385 and :1416-1419"""
    avg = np.exp(log_avg - np.amax(log_avg))
    avg /= np.sum(avg.reshape(len(avg), -1), axis=1).reshape((-1,) + (1,) * len(pb.shape))
    means = np.array([[np.sum(p * g) for p in avg] for g in pb.grid])
    return avg, means


def online_steps(pb, ops, rows, use_scipy=True):
    """OnlineStudy.step for the hypotheses `rows` over the whole series (core.py:2143-2175): ONE likelihood per step
    shared by all hypotheses, then per hypothesis transition (t = -1, core.py:2167) x likelihood, sum, divide.
    Returns the log-evidences of the hypotheses."""
    post = {b: None for b in rows}
    logE = {b: np.log(pb.lc) for b in rows}
    for i in range(pb.T):
        lik = pb.likelihood(i)
        for b in rows:
            alpha = (pb.prior if post[b] is None else apply_ops(pb, ops, b, -1, False, post[b], use_scipy)) * lik
            ni = np.sum(alpha)
            logE[b] += np.log(ni)
            post[b] = alpha / ni
    return np.array([logE[b] for b in rows])


def cell_updates(pb, n_combos, forward_only=False, evidence_only=False):
    per_pass = n_combos * pb.T * int(np.prod(pb.shape))
    return per_pass if (forward_only or evidence_only) else 2 * per_pass
