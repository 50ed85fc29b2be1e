"""Transition models (parameter-dynamics plugins) of the B200 engine.

Public surface mirrors bayesloop/transitionModels.py: every model carries `study`, `latticeConstant`,
`hyperParameterNames`, `hyperParameterValues`, `prior` (hyper-prior), `tOffset`, composites carry `models`, and
`str(model)` yields the same load-bearing labels ('Change-point', 'Break-point', 'Serial transition model', ...).

The reference applies a model by calling `computeForwardPrior(posterior, t)` / `computeBackwardPrior(posterior, t)`
on a NumPy grid at every time step.  Here a model is LOWERED instead: `lower(ctx, window)` appends flat operators
(include/blgrid.h `enum blg_op_kind`) to a per-combo transition program -- operator kind, grid axis, a parameter
per hyper-parameter combination, and the half-open range of time-step indices in which the operator is active in
the forward and in the backward pass.  The CUDA kernels then run the whole program inside their time loop.

    GaussianRandomWalk   -> GRW(axis, sigma/latticeConstant[axis], radius)   transitionModels.py:96-115
    RegimeSwitch         -> REGIME(10**log10pMin * prod(latticeConstant))     transitionModels.py:394-412
    ChangePoint          -> RESET(prod(latticeConstant)) at ONE step          transitionModels.py:289-314
    Independent          -> RESET(1.0) at every step                          transitionModels.py:339-360
    NotEqual             -> NOTEQUAL(limit)                                   transitionModels.py:450-471
    Static               -> (nothing)                                         transitionModels.py:49-60
    Combined             -> children in listed order, same window             transitionModels.py:632-662
    Serial               -> child k limited to the steps whose time stamp lies between break k and k+1,
                            embedded change-points appended as RESETs         transitionModels.py:756-818

A transition model written as arbitrary Python (a subclass that only overrides computeForwardPrior) cannot run on
the device; `lower` raises NotImplementedError naming it -- there is no CPU fallback.
"""
from collections.abc import Iterable

import numpy as np

from .exceptions import ConfigurationError

# op-codes: keep in sync with include/blgrid.h (enum blg_op_kind)
OP_GRW = 1
OP_REGIME = 2
OP_RESET = 3
OP_NOTEQUAL = 4

ALWAYS = (-2 ** 31, 2 ** 31 - 1)


class Window:
    """Per-combo active ranges of time-step indices: forward [f_lo, f_hi), backward [b_lo, b_hi)."""

    def __init__(self, f_lo, f_hi, b_lo, b_hi):
        self.f_lo, self.f_hi, self.b_lo, self.b_hi = f_lo, f_hi, b_lo, b_hi

    @staticmethod
    def everything(B):
        lo = np.full(B, ALWAYS[0], dtype=np.int64)
        hi = np.full(B, ALWAYS[1], dtype=np.int64)
        return Window(lo, hi, lo.copy(), hi.copy())

    def clip(self, f_lo, f_hi, b_lo, b_hi):
        return Window(np.maximum(self.f_lo, f_lo), np.minimum(self.f_hi, f_hi),
                      np.maximum(self.b_lo, b_lo), np.minimum(self.b_hi, b_hi))


class LoweringContext:
    """Everything a model needs to emit its operators for B hyper-parameter combinations at once.

    hyper:       float array [B, H]; column order = flattened hyper-parameter order (core.py:634-647)
    timestamps:  float array [T] of formattedTimestamps -- the values handed to the models as `t` (core.py:411)
    online:      OnlineStudy hands t = -1 at every step (core.py:2166-2167): windows collapse to
                 always/never
    """

    def __init__(self, parameterNames, latticeConstant, hyper, timestamps, online=False):
        self.parameterNames = list(parameterNames)
        self.lattice = [float(x) for x in latticeConstant]
        self.lcProd = float(np.prod(self.lattice))
        self.hyper = np.asarray(hyper, dtype=float)
        self.B = self.hyper.shape[0]
        self.online = online
        ts = np.asarray(timestamps, dtype=float)
        if not online and ts.size > 1 and np.any(np.diff(ts) <= 0):
            raise NotImplementedError('time stamps must be strictly increasing for time-dependent transition models')
        self.ts = ts
        self.tsBack = ts - 1.  # computeBackwardPrior(p, t) = computeForwardPrior(p, t - 1)
        self.ops = []
        self.usesReset = False

    # -- index helpers -------------------------------------------------------------------------------------
    def _first_at_or_after(self, values, backward):
        base = self.tsBack if backward else self.ts
        return np.searchsorted(base, values, side='left').astype(np.int64)

    def steps_equal(self, values):
        """Per combo: the single step index whose time stamp equals `values` (forward) / equals values+1
        (backward), as half-open ranges; empty range when no stamp matches."""
        out = []
        for backward in (False, True):
            base = self.tsBack if backward else self.ts
            pos = self._first_at_or_after(values, backward)
            safe = np.minimum(pos, base.size - 1)
            hit = (pos < base.size) & (base[safe] == values)
            out += [np.where(hit, pos, 0), np.where(hit, pos + 1, 0)]
        return out

    def steps_between(self, lower, upper):
        """Per combo: steps whose time stamp t satisfies lower <= t < upper (forward; backward uses t - 1)."""
        out = []
        for backward in (False, True):
            out += [self._first_at_or_after(lower, backward), self._first_at_or_after(upper, backward)]
        return out

    def emit(self, kind, axis, param, radius, window):
        f_lo, f_hi, b_lo, b_hi = window.f_lo, window.f_hi, window.b_lo, window.b_hi
        if self.online:  # t = -1 for ever: an operator is either always or never applied
            on = (f_lo <= 0) & (f_hi >= 1)
            f_lo = np.where(on, ALWAYS[0], 0)
            f_hi = np.where(on, ALWAYS[1], 0)
            b_lo, b_hi = f_lo, f_hi
        clipInt = lambda a: np.clip(a, ALWAYS[0], ALWAYS[1]).astype(np.int32)
        self.ops.append(dict(kind=kind, axis=axis,
                             param=np.broadcast_to(np.asarray(param, dtype=float), (self.B,)).copy(),
                             radius=np.broadcast_to(np.asarray(radius, dtype=np.int32), (self.B,)).copy(),
                             window=np.stack([clipInt(f_lo), clipInt(f_hi), clipInt(b_lo), clipInt(b_hi)], axis=1)))
        if kind == OP_RESET:
            self.usesReset = True


def assign_columns(model, start=0):
    """Give every node its slice of the flattened hyper-parameter vector: sub-models first (depth-first), then the
    node's own names (the order of Study._unpackHyperParameters, core.py:634-647).  Returns the next free column."""
    pos = start
    for sub in getattr(model, 'models', []):
        pos = assign_columns(sub, pos)
    count = len(getattr(model, 'hyperParameterNames', []))
    model._columns = list(range(pos, pos + count))
    return pos + count


class TransitionModel:
    """Base class of all transition models."""

    def lower(self, ctx, window):
        raise NotImplementedError('Transition model "{}" has no device lowering; arbitrary Python dynamics cannot '
                                  'run inside the CUDA time loop (no CPU fallback).'.format(self))

    def computeForwardPrior(self, posterior, t):
        raise NotImplementedError('bayesloop_b200 applies transition models on the device; see lower().')

    def computeBackwardPrior(self, posterior, t):
        return self.computeForwardPrior(posterior, t - 1)

    def _init_common(self):
        self.study = None
        self.latticeConstant = None
        self.tOffset = 0

    # The back-pointer `study` (core.py:281 of the reference) is a weak proxy here; a proxy forwards __reduce_ex__
    # to the study it points to, so pickle / cloudpickle would serialise a SECOND, half-initialised study in its
    # place.  It is dropped on the way out and re-attached by Study.__setstate__.
    def __getstate__(self):
        d = self.__dict__.copy()
        d['study'] = None
        return d


def _as_array(value):
    return np.array(value) if isinstance(value, (list, tuple)) else value


class Static(TransitionModel):
    """Parameters do not change."""

    def __init__(self):
        self._init_common()
        self.hyperParameterNames = []
        self.hyperParameterValues = []
        self.prior = None

    def __str__(self):
        return 'Static/constant parameter values'

    def lower(self, ctx, window):
        return


class GaussianRandomWalk(TransitionModel):
    """Gaussian fluctuations of ONE parameter (`target`) with standard deviation `value`."""

    def __init__(self, name='sigma', value=None, target=None, prior=None):
        self._init_common()
        self.hyperParameterNames = [name]
        self.hyperParameterValues = [_as_array(value)]
        self.prior = prior
        self.selectedParameter = target
        if target is None:
            raise ConfigurationError('No parameter set for transition model "GaussianRandomWalk"')

    def __str__(self):
        return 'Gaussian random walk'

    def lower(self, ctx, window):
        if self.selectedParameter not in ctx.parameterNames:
            raise ConfigurationError('Gaussian random walk targets unknown parameter "{}".'
                                     .format(self.selectedParameter))
        axis = ctx.parameterNames.index(self.selectedParameter)
        sigma = ctx.hyper[:, self._columns[0]] / ctx.lattice[axis]
        # radius of scipy.ndimage.gaussian_filter1d: int(truncate * sigma + 0.5), truncate = 4.0; the filter is
        # skipped for sigma <= 0 (transitionModels.py:110-113)
        radius = np.where(sigma > 0., np.floor(4.0 * np.maximum(sigma, 0.) + 0.5), 0).astype(np.int64)
        ctx.emit(OP_GRW, axis, sigma, radius, window)


class ChangePoint(TransitionModel):
    """Parameter distribution is reset to the prior right after time stamp `value`."""

    def __init__(self, name='tChange', value=None, prior=None):
        self._init_common()
        self.hyperParameterNames = [name]
        self.hyperParameterValues = [_as_array(value)]
        self.prior = prior

    def __str__(self):
        return 'Change-point'

    def lower(self, ctx, window):
        at = ctx.hyper[:, self._columns[0]]
        ctx.emit(OP_RESET, 0, ctx.lcProd, 0, window.clip(*ctx.steps_equal(at)))


class Independent(TransitionModel):
    """Every observation starts from the prior again."""

    def __init__(self):
        self._init_common()
        self.hyperParameterNames = []
        self.hyperParameterValues = []
        self.prior = None

    def __str__(self):
        return 'Independent observations model'

    def lower(self, ctx, window):
        ctx.emit(OP_RESET, 0, 1.0, 0, window)


class RegimeSwitch(TransitionModel):
    """Minimal probability density 10**value for every parameter value at every step."""

    def __init__(self, name='log10pMin', value=None, prior=None):
        self._init_common()
        self.hyperParameterNames = [name]
        self.hyperParameterValues = [_as_array(value)]
        self.prior = prior

    def __str__(self):
        return 'Regime-switching model'

    def lower(self, ctx, window):
        limit = (10. ** ctx.hyper[:, self._columns[0]]) * ctx.lcProd
        ctx.emit(OP_REGIME, 0, limit, 0, window)


class NotEqual(TransitionModel):
    """Inverted parameter distribution with minimal probability 10**value."""

    def __init__(self, name='log10pMin', value=None, prior=None):
        self._init_common()
        self.hyperParameterNames = [name]
        self.hyperParameterValues = [_as_array(value)]
        self.prior = prior

    def __str__(self):
        return 'Not-Equal model'

    def lower(self, ctx, window):
        limit = (10. ** ctx.hyper[:, self._columns[0]]) * ctx.lcProd
        ctx.emit(OP_NOTEQUAL, 0, limit, 0, window)


class CombinedTransitionModel(TransitionModel):
    """Several models act at every time step, in the listed order."""

    def __init__(self, *args):
        self._init_common()
        self.models = args
        if any(str(a) == 'Break-point' for a in args):
            raise ConfigurationError('The "BreakPoint" transition model can only be used with the '
                                     '"SerialTransitionModel" class.')

    def __str__(self):
        return 'Combined transition model'

    def lower(self, ctx, window):
        for sub in self.models:
            sub.lower(ctx, window)


class BreakPoint(TransitionModel):
    """Marker for a structural break inside a SerialTransitionModel."""

    def __init__(self, name='tBreak', value=None, prior=None):
        self.name = name
        self.value = _as_array(value)
        self.prior = prior

    def __str__(self):
        return 'Break-point'


class SerialTransitionModel(TransitionModel):
    """Different models act in consecutive time intervals separated by break-points (dynamics change) or
    change-points (dynamics change AND the parameter distribution is reset)."""

    def __init__(self, *args):
        self._init_common()
        self.hyperParameterNames = []
        self.hyperParameterValues = []
        self.prior = []
        self.models = []
        mask = []
        for arg in args:
            label = str(arg)
            if label == 'Break-point':
                name, value, prior, isChange = arg.name, arg.value, arg.prior, 0
            elif label == 'Change-point':
                name, value, prior, isChange = arg.hyperParameterNames[0], arg.hyperParameterValues[0], arg.prior, 1
            else:
                self.models.append(arg)
                continue
            if not (isinstance(value, str) and value == 'all') and isinstance(value, Iterable):
                value = np.array(value)
            self.hyperParameterNames.append(name)
            self.hyperParameterValues.append(value)
            self.prior.append(prior)
            mask.append(isChange)
        self.changePointMask = np.array(mask).astype(bool)

        firsts = []
        for v in self.hyperParameterValues:
            if isinstance(v, str):
                firsts.append(None)
            elif isinstance(v, Iterable):
                firsts.append(v[0])
            else:
                firsts.append(v)
        for a, b in zip(firsts, firsts[1:]):
            if a is not None and b is not None and not a < b:
                raise ConfigurationError('Time steps for structural breaks and/or change-points have to be passed in '
                                         'monotonically increasing order.')
        if len(self.models) - 1 != len(self.hyperParameterValues):
            raise ConfigurationError('Wrong number of structural breaks/change-points and models. For n models, n-1 '
                                     'structural breaks/change-points are required.')

    def __str__(self):
        return 'Serial transition model'

    def lower(self, ctx, window):
        B = ctx.B
        points = ctx.hyper[:, self._columns] if self._columns else np.empty((B, 0))
        # model k is active while  #(points <= t) == k  (transitionModels.py:768); with the points of a combo
        # sorted that is  sorted[k-1] <= t < sorted[k]
        ordered = np.sort(points, axis=1)
        edges = np.concatenate([np.full((B, 1), -np.inf), ordered, np.full((B, 1), np.inf)], axis=1)
        for k, sub in enumerate(self.models):
            sub.lower(ctx, window.clip(*ctx.steps_between(edges[:, k], edges[:, k + 1])))
        for j in np.flatnonzero(self.changePointMask):  # embedded change-points (transitionModels.py:788-818)
            ctx.emit(OP_RESET, 0, ctx.lcProd, 0, window.clip(*ctx.steps_equal(points[:, j])))


class _Unsupported(TransitionModel):
    """Transition models of the reference that have no device operator in this engine (each needs its own kernel
    family: 3x-padded FFT convolution, cubic-spline shift, dense 2-D convolution; SURVEY.md section 2 row 8 marks them
    out of scope).  Constructing one fails immediately and says so, instead of an AttributeError on `bl.tm.<name>`
    or a late failure inside fit()."""
    label = ''
    reference = ''

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('bayesloop_b200 has no device kernel for the transition model "{}" ({}); there is '
                                  'no CPU fallback.  Supported: Static, GaussianRandomWalk, ChangePoint, RegimeSwitch, '
                                  'Independent, NotEqual, CombinedTransitionModel, SerialTransitionModel, BreakPoint.'
                                  .format(self.label, self.reference))


class AlphaStableRandomWalk(_Unsupported):
    label, reference = 'AlphaStableRandomWalk', 'bayesloop/transitionModels.py:121-260'


class Deterministic(_Unsupported):
    label, reference = 'Deterministic', 'bayesloop/transitionModels.py:477-606'


class BivariateRandomWalk(_Unsupported):
    label, reference = 'BivariateRandomWalk', 'bayesloop/transitionModels.py:843-911'
