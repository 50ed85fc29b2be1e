"""Observation models (likelihood plugins) of the B200 engine.

Public surface mirrors bayesloop/observationModels.py: every model carries `name`, `segmentLength`,
`parameterNames`, `parameterValues`, `multiplyLikelihoods`, `prior`, `estimateParameterValues` and (where the
reference has one) a `jeffreys` prior.  What differs is WHERE the likelihood is evaluated: the closed-form models
do not implement a NumPy `pdf`; they name a device op-code (`deviceKind`, include/blgrid.h `enum blg_om_kind`) and
the fused CUDA kernels evaluate the density over the parameter grid inside the forward/backward time loop
(bayesloop_b200/csrc/blgrid_device.cuh).  Models without an op-code -- user subclasses that implement `pdf`, and
the `NumPy`/`SciPy`/`SymPy` wrappers -- are honoured through a likelihood table [T x G] that is filled ONCE per
fit on the host by `processedPdf` (it does not depend on the hyper-parameter combination) and streamed by the
kernels (`BLG_OM_TABLE`).
"""
import inspect
import warnings

import numpy as np

from .exceptions import ConfigurationError
from .helper import cint, oint

# op-codes: keep in sync with include/blgrid.h (enum blg_om_kind)
KIND_POISSON = 1
KIND_GAUSSIAN = 2
KIND_SCALED_AR1 = 3
KIND_AR1 = 4
KIND_WHITE_NOISE = 5
KIND_GAUSSIAN_MEAN = 6
KIND_LAPLACE = 7
KIND_BERNOULLI = 8
KIND_TABLE = 100


class ObservationModel:
    """Base class; subclass it and implement `pdf(grid, dataSegment)` to plug a custom likelihood in."""

    deviceKind = KIND_TABLE  # no device op-code -> host-precomputed likelihood table
    name = 'observation model'
    segmentLength = 1
    multiplyLikelihoods = True
    prior = None

    def __str__(self):
        return self.name

    def processedPdf(self, grid, dataSegment):
        """Likelihood of one data segment on the grid, honouring missing values and multi-column data.

        Only used to fill the likelihood table of models WITHOUT a device op-code (contract of
        bayesloop/observationModels.py:35-56): NaN anywhere in the segment -> array of ones; 2-D segments of a
        model with `multiplyLikelihoods` -> product of the per-column likelihoods.
        """
        segment = np.asarray(dataSegment)
        if segment.ndim == 2 and self.multiplyLikelihoods:
            out = np.ones_like(grid[0], dtype=float)
            for column in segment.T:
                out = out * self.processedPdf(grid, column)
            return out
        if np.isnan(segment.astype(float)).any():
            return np.ones_like(grid[0], dtype=float)
        return self.pdf(grid, segment)

    def pdf(self, grid, dataSegment):
        raise NotImplementedError('{}: closed-form models are evaluated on the device (op-code {}); custom models '
                                  'must implement pdf(grid, dataSegment).'.format(self.name, self.deviceKind))

    def estimateParameterValues(self, name, rawData):
        raise ConfigurationError('{} cannot estimate values for parameter "{}".'.format(self.name, name))


def _pick_prior(model, prior):
    if isinstance(prior, str) and prior == 'Jeffreys':
        return model.jeffreys
    return prior


def _spread(rawData):
    flat = np.ravel(np.asarray(rawData, dtype=float))
    return np.nanmean(flat), np.nanstd(flat)


class Poisson(ObservationModel):
    """Poisson counts with rate lambda (device op-code POISSON; reference pdf observationModels.py:502)."""
    deviceKind = KIND_POISSON

    def __init__(self, name='lambda', value=None, prior='Jeffreys'):
        self.name = 'Poisson'
        self.segmentLength = 1
        self.parameterNames = [name]
        self.parameterValues = [value]
        self.multiplyLikelihoods = True
        self.prior = _pick_prior(self, prior)

    def estimateParameterValues(self, name, rawData):
        if name != self.parameterNames[0]:
            raise ConfigurationError('Poisson model does not contain a parameter "{}".'.format(name))
        top = 1.25 * np.nanmax(np.ravel(np.asarray(rawData, dtype=float)))
        return oint(0, top, 1000)

    def jeffreys(self, x):
        return np.sqrt(1. / x)


class Gaussian(ObservationModel):
    """Independent normal observations, parameters (mean, std) (op-code GAUSSIAN; observationModels.py:566-567)."""
    deviceKind = KIND_GAUSSIAN

    def __init__(self, name1='mean', value1=None, name2='std', value2=None, prior='Jeffreys'):
        self.name = 'Gaussian observations'
        self.segmentLength = 1
        self.parameterNames = [name1, name2]
        self.parameterValues = [value1, value2]
        self.multiplyLikelihoods = True
        self.prior = _pick_prior(self, prior)

    def estimateParameterValues(self, name, rawData):
        mean, std = _spread(rawData)
        if name == self.parameterNames[0]:
            return cint(mean - 2 * std, mean + 2 * std, 200)
        if name == self.parameterNames[1]:
            return oint(0, 2 * std, 200)
        raise ConfigurationError('Gaussian model does not contain a parameter "{}".'.format(name))

    def jeffreys(self, mu, sigma):
        return 1. / sigma ** 2.


class Laplace(ObservationModel):
    """Double-exponential observations, parameters (mean, scale) (op-code LAPLACE; observationModels.py:635)."""
    deviceKind = KIND_LAPLACE

    def __init__(self, name1='mean', value1=None, name2='scale', value2=None, prior='Jeffreys'):
        self.name = 'Laplace observations'
        self.segmentLength = 1
        self.parameterNames = [name1, name2]
        self.parameterValues = [value1, value2]
        self.multiplyLikelihoods = True
        self.prior = _pick_prior(self, prior)

    def estimateParameterValues(self, name, rawData):
        mean, std = _spread(rawData)
        if name == self.parameterNames[0]:
            return cint(mean - 2 * std, mean + 2 * std, 200)
        if name == self.parameterNames[1]:
            return oint(0, np.sqrt(2) * std, 200)
        raise ConfigurationError('Laplace model does not contain a parameter "{}".'.format(name))

    def jeffreys(self, mu, scale):
        return 1. / scale ** 2.


class GaussianMean(ObservationModel):
    """Normal observations with a known per-step error: data rows are [value, std] (op-code GAUSSIAN_MEAN;
    observationModels.py:705-706).  The two data columns form ONE observation, they are not multiplied."""
    deviceKind = KIND_GAUSSIAN_MEAN

    def __init__(self, name='mean', value=None, prior=None):
        self.name = 'Gaussian mean model'
        self.segmentLength = 1
        self.parameterNames = [name]
        self.parameterValues = [value]
        self.multiplyLikelihoods = False
        self.prior = prior

    def estimateParameterValues(self, name, rawData):
        if name != self.parameterNames[0]:
            raise ConfigurationError('Gaussian mean model does not contain a parameter "{}".'.format(name))
        obs = np.asarray(rawData, dtype=float)[:, 0]
        lo, hi = np.nanmin(obs), np.nanmax(obs)
        return oint(lo - (hi - lo), hi + (hi - lo), 1000)


class WhiteNoise(ObservationModel):
    """Zero-mean normal observations with amplitude std (op-code WHITE_NOISE; observationModels.py:767)."""
    deviceKind = KIND_WHITE_NOISE

    def __init__(self, name='std', value=None, prior='Jeffreys'):
        self.name = 'White noise process (Zero-mean Gaussian)'
        self.segmentLength = 1
        self.parameterNames = [name]
        self.parameterValues = [value]
        self.multiplyLikelihoods = True
        self.prior = _pick_prior(self, prior)

    def estimateParameterValues(self, name, rawData):
        if name != self.parameterNames[0]:
            raise ConfigurationError('White noise model does not contain a parameter "{}".'.format(name))
        return oint(0, 2 * _spread(rawData)[1], 1000)

    def jeffreys(self, sigma):
        return 1. / sigma


class Bernoulli(ObservationModel):
    """Success probability p of 0/1 observations (op-code BERNOULLI; observationModels.py:430-439)."""
    deviceKind = KIND_BERNOULLI

    def __init__(self, name='p', value=None, prior='Jeffreys'):
        self.name = 'Bernoulli'
        self.segmentLength = 1
        self.parameterNames = [name]
        self.parameterValues = [value]
        self.multiplyLikelihoods = True
        self.prior = _pick_prior(self, prior)

    def estimateParameterValues(self, name, rawData):
        if name != self.parameterNames[0]:
            raise ConfigurationError('Bernoulli model does not contain a parameter "{}".'.format(name))
        return cint(0, 1, 1000)

    def jeffreys(self, x):
        return 1. / np.sqrt(x * (1. - x))


class AR1(ObservationModel):
    """d_t = r d_{t-1} + s e_t, parameters (r, s), two-point segments (op-code AR1; observationModels.py:830-831)."""
    deviceKind = KIND_AR1

    def __init__(self, name1='correlation coefficient', value1=None, name2='noise amplitude', value2=None,
                 prior=None):
        self.name = 'Autoregressive process of first order (AR1)'
        self.segmentLength = 2
        self.parameterNames = [name1, name2]
        self.parameterValues = [value1, value2]
        self.multiplyLikelihoods = True
        self.prior = prior

    def estimateParameterValues(self, name, rawData):
        if name == self.parameterNames[0]:
            return oint(-1, 1, 200)
        if name == self.parameterNames[1]:
            return oint(0, 2 * _spread(rawData)[1], 200)
        raise ConfigurationError('AR1 model does not contain a parameter "{}".'.format(name))


class ScaledAR1(AR1):
    """d_t = r d_{t-1} + s sqrt(1-r^2) e_t, parameters (r, s) (op-code SCALED_AR1; observationModels.py:892-896)."""
    deviceKind = KIND_SCALED_AR1

    def __init__(self, name1='correlation coefficient', value1=None, name2='standard deviation', value2=None,
                 prior=None):
        AR1.__init__(self, name1, value1, name2, value2, prior)
        self.name = 'Scaled autoregressive process of first order (AR1)'


# --------------------------------------------------------------------------------------- table-path wrappers
def _name_value_pairs(args):
    if len(args) == 1 and isinstance(args[0], dict):
        warnings.warn('Passing parameters as a dictionary is deprecated; pass names and values alternately.',
                      DeprecationWarning)
        return list(args[0].keys()), list(args[0].values())
    return list(args[::2]), list(args[1::2])


class NumPy(ObservationModel):
    """Likelihood given as a Python function `f(data, *parameterArrays)` (reference: observationModels.py:59-143).
    Evaluated once per time step on the host into the likelihood table; no device op-code."""

    def __init__(self, function, *args, **kwargs):
        if not callable(function):
            raise ConfigurationError('Expected a function as the first argument of NumPy observation model')
        unknown = set(kwargs) - {'prior'}
        if unknown:
            raise TypeError("__init__() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
        self.function = function
        self.name = function.__name__
        self.segmentLength = 1
        self.multiplyLikelihoods = False
        self.parameterNames, self.parameterValues = _name_value_pairs(args)
        argNames = list(inspect.signature(function).parameters)
        if len(argNames) - 1 != len(self.parameterNames):
            raise ConfigurationError('Supplied function has {} parameters, observation model has {}'
                                     .format(len(argNames) - 1, len(self.parameterNames)))
        if argNames[0] != 'data':
            raise ConfigurationError('First argument of supplied function must be called "data"')
        self.prior = kwargs.get('prior', None)

    def pdf(self, grid, dataSegment):
        return self.function(dataSegment[0], *grid)


class SciPy(ObservationModel):
    """Likelihood from a scipy.stats distribution (reference: observationModels.py:146-269); table path."""

    def __init__(self, rv, *args, **kwargs):
        module = getattr(rv, '__module__', '') or ''
        if not module.startswith('scipy.stats'):
            raise ConfigurationError('SciPy observation model must contain SciPy probability distribution')
        unknown = set(kwargs) - {'prior', 'fixedParameters'}
        if unknown:
            raise TypeError("__init__() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
        self.rv = rv
        self.name = rv.name
        self.parameterNames, self.parameterValues = _name_value_pairs(args)
        self.prior = kwargs.get('prior', None)
        self.fixedParameterDict = kwargs.get('fixedParameters', {})
        self.segmentLength = 1
        self.multiplyLikelihoods = True
        self.isContinuous = hasattr(rv, 'pdf')
        shapes = rv.shapes.split(', ') if rv.shapes else []
        shapes.append('loc')
        if self.isContinuous:
            shapes.append('scale')
        free = [s for s in shapes if s not in self.fixedParameterDict]
        if len(self.parameterNames) == 0:
            self.parameterNames, self.parameterValues = free, [None] * len(free)
        stray = set(self.parameterNames) - set(free)
        if stray:
            raise ConfigurationError('The following parameter names from the observation model do not match the '
                                     'parameter names of the SciPy distribution: {} (options: {})'
                                     .format(sorted(stray), free))

    def pdf(self, grid, dataSegment):
        params = dict(zip(self.parameterNames, grid))
        params.update(self.fixedParameterDict)
        density = self.rv.pdf if self.isContinuous else self.rv.pmf
        return density(dataSegment[0], **params)


class SymPy(ObservationModel):
    """Likelihood from a sympy.stats random variable (reference: observationModels.py:272-391); table path.
    As in the reference the Jeffreys-prior attempt never succeeds, so the prior is flat
    unless given."""

    def __init__(self, rv, *args, **kwargs):
        import sympy.abc
        from scipy.special import factorial, iv
        from sympy import lambdify
        from sympy.stats import density
        module = getattr(rv, '__module__', '') or ''
        if not module.startswith('sympy.stats'):
            raise ConfigurationError('SymPy observation model must contain SymPy random variable.')
        unknown = set(kwargs) - {'prior', 'determineJeffreysPrior'}
        if unknown:
            raise TypeError("__init__() got an unexpected keyword argument '{}'".format(sorted(unknown)[0]))
        self.name = str(rv)
        self.parameterNames, self.parameterValues = _name_value_pairs(args)
        symbols = _free_symbols(rv)
        names = [str(s) for s in symbols]
        if len(self.parameterNames) == 0:
            self.parameterNames, self.parameterValues = names, [None] * len(names)
        stray = set(self.parameterNames) - set(names)
        if stray:
            raise ConfigurationError('The following parameter names from the observation model do not match the '
                                     'names of SymPy random variables: {}'.format(sorted(stray)))
        ordered = [symbols[names.index(n)] for n in self.parameterNames]
        self.prior = kwargs.get('prior', None)
        self.segmentLength = 1
        self.multiplyLikelihoods = True
        x = sympy.abc.x
        self.density = lambdify([x] + ordered, density(rv)(x),
                                modules=['numpy', {'factorial': factorial, 'besseli': iv}])

    def pdf(self, grid, dataSegment):
        return self.density(dataSegment[0], *grid)


def _free_symbols(rv):
    for arg in rv._sorted_args:
        dist = getattr(arg, 'distribution', None)
        if dist is not None:
            return list(dist.free_symbols)
    return []
