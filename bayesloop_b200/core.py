"""Study classes of the B200 engine: same public API as bayesloop/core.py, different machine underneath.

The reference runs `Study.fit` as a Python loop over time steps (core.py:372-470) and `HyperStudy.fit` as a Python
loop over hyper-parameter combinations that calls it (core.py:1349-1366).  Here the combinations are the BATCH: the
transition-model tree is lowered once into a flat per-combo program (transitionModels.lower), and one
`blg_forward` + one `blg_backward` call (include/blgrid.h) run the whole time loop for a wave of combinations
inside persistent sm_100a kernels.  Host code below only prepares inputs (grid, prior, hyper-grid, windows) and
post-processes O(B) / O(T x G) results; nothing here touches a grid cell per time step.
"""
import math
from collections.abc import Iterable

import weakref

import numpy as np
import torch

from . import engine as _engine
from . import transitionModels as _tm
from .exceptions import ConfigurationError, PostProcessingError
from .helper import assignNestedItem, flatten, is_regular, recursiveIndex
from .observationModels import KIND_GAUSSIAN_MEAN, KIND_POISSON, KIND_TABLE, ObservationModel
from .preprocessing import movingWindow
from .transitionModels import TransitionModel

COAL_MINING_DISASTERS = (  # UK coal mining disasters per year, 1852-1961 (the series of core.py:82-86)
    "5410434063340263354531441553425223421322111130010110031032201110"
    "1010002100011023311211112330001400010000010010")


def _is_sympy_rv(obj):
    try:
        import sympy.stats.rv
        return isinstance(obj, sympy.stats.rv.RandomSymbol)
    except Exception:
        return False


def _sympy_density(rv, symbol):
    from scipy.special import beta as beta_func
    from scipy.special import factorial
    from sympy import lambdify
    from sympy.stats import density
    return lambdify([symbol], density(rv)(symbol), modules=['numpy', {'factorial': factorial, 'beta': beta_func}])


def _free_symbols(rv):
    for arg in rv._sorted_args:
        dist = getattr(arg, 'distribution', None)
        if dist is not None:
            return list(dist.free_symbols)
    return []


def _share_structure(ops, T):
    """Change-point prefix sharing: can the combinations of a lowered sweep be arranged as
    (groups) x (change-points), where the members of a group differ ONLY in the step at which ONE reset operator fires?
    A ChangePoint erases the history (transitionModels.py:300-312): before it every member of the group filters
    exactly like the change-point-free model, after it the backward message is that model's as well.
    Returns None or dict(k, group[B], cidx[B], cvals[nC], nG, nC)."""
    resets = [k for k, op in enumerate(ops) if op['kind'] == _tm.OP_RESET]
    if len(resets) != 1 or T < 3:
        return None
    k = resets[0]
    win = np.asarray(ops[k]['window'], dtype=np.int64)
    B = len(win)
    if B < 4:
        return None
    c = win[:, 0]
    if not (np.all(win[:, 1] == c + 1) and np.all(win[:, 2] == c + 1) and np.all(win[:, 3] == c + 2)):
        return None  # not a single-step reset (e.g. Independent: every step)
    if c.min() < 0 or c.max() > T - 2:
        return None
    keys = []
    for j, op in enumerate(ops):
        w = np.asarray(op['window'])
        if j != k and not np.all(w == w[0]):
            return None  # another operator's activity depends on the combination (Serial models): not shareable
        if j != k and not (np.all(w[:, 0] <= 0) and np.all(w[:, 1] >= T) and np.all(w[:, 2] <= 0) and np.all(w[:, 3] >= T)):
            # an operator that is active on a RANGE of steps only (a segment of a Serial model next to the change-point):
            # the shared schedule runs the passes on windows of the sequence, whose step indices start at 0 again
            return None
        keys.append(np.asarray(op['param'], dtype=float).reshape(B, 1))
        keys.append(np.asarray(op['radius'], dtype=float).reshape(B, 1))
    _, group = np.unique(np.hstack(keys), axis=0, return_inverse=True)
    group = np.asarray(group).reshape(-1)
    cvals, cidx = np.unique(c, return_inverse=True)
    cidx = np.asarray(cidx).reshape(-1)
    nG, nC = int(group.max()) + 1, len(cvals)
    if nC < 2 or nG * nC != B or len(np.unique(group * nC + cidx)) != B:
        return None  # not a full (groups x change-points) rectangle
    return dict(k=k, group=group, cidx=cidx, cvals=cvals.astype(np.int64), nG=nG, nC=nC)


def _combo_cost(ops, B, T):
    """Predicted cost of one time step of every combination, in units of one convolution tap per cell: the elementwise
    work of a step (~18: DESIGN.md 6 counts 46 FMA per update at 28 taps) plus 2 R + 1 taps per random-walk operator,
    weighted with the share of the steps on which the operator is active (segments of Serial models)."""
    cost = np.full(int(B), 18.0)
    for op in ops:
        radius = np.asarray(op['radius'], dtype=np.int64).reshape(int(B))
        w = np.asarray(op['window'], dtype=np.int64).reshape(int(B), 4)
        active = np.clip(np.minimum(w[:, 1], T) - np.maximum(w[:, 0], 0), 0, T) / float(max(T, 1))
        cost += np.where(radius > 0, 2 * radius + 1, 0) * active
    return cost


def _deal_shared(share, ops, size, slots, fits_all_groups):
    """How a (groups) x (change-points) sweep is dealt over `size` ranks: 'groups' (whole groups, every change-point),
    'changepoints' (every group, a subset of the change-points) or None (no sharing: rows dealt by predicted cost).  The
    answer depends only on its arguments, which are the same on every rank.

    A shared sweep launches one forward and one backward pass per change-point with B = (groups of the rank) combos
    (HyperStudy._executeSharedSweep), plus two full-length passes of the change-point-free run of every group of the
    rank.  With `slots` combos resident on the device at a time a launch takes the makespan of its combos on
    the slots: a launch with few combos is as long as its costliest combo and leaves slots idle.  Dealing GROUPS
    shrinks every launch (C4 on 8 GPUs: 12-13 combos on 15 resident clusters, the launches as long as the widest
    random walk: 5.5 x one GPU measured, profiles/r2Y2_bench_n8.json); dealing CHANGE-POINTS keeps the launches of the
    single-GPU run and the ranks balanced ((T - c) + (c + 2) executed steps per combo whatever c), at the price of
    every rank repeating the change-point-free runs (2 of 2 + nC/size pass-equivalents) and holding the two shared
    sequences of ALL groups (`fits_all_groups`)."""
    nG, nC = share['nG'], share['nC']
    rep = np.full(nG, -1, dtype=np.int64)
    rep[share['group']] = np.arange(len(share['group']))  # one representative combo per group
    cost = _combo_cost(ops, len(share['group']), 1 << 30)[rep]  # (the operators next to a shared reset are always active)

    def launch(sel):  # makespan of one launch: combos by descending cost onto the least loaded of `slots` slots
        c = np.sort(cost[sel])[::-1]
        if len(c) <= slots:
            return float(c[0])
        load = np.zeros(int(slots))
        for x in c:
            load[np.argmin(load)] += x
        return float(load.max())

    cand = {}
    if nC // size >= 2 and (fits_all_groups or size == 1):
        cand['changepoints'] = launch(np.arange(nG)) * (2 + -(-nC // size))
    if nG >= size:
        cand['groups'] = max(launch(np.arange(r, nG, size)) for r in range(size)) * (2 + nC)
    if size == 1:
        return 'groups' if cand else None  # one rank: the two deals are the same thing
    if not cand:
        return None
    return min(sorted(cand), key=cand.get)  # ties: 'changepoints' (balanced whatever the groups cost)


class _Session:
    """Device-side constants of one fit: plan, data series, prior, likelihood table."""

    def __init__(self, study, eng):
        om = study.observationModel
        self.eng = eng
        self.T = len(study.formattedData)
        self.gridSize = list(study.gridSize)
        self.G = int(np.prod(self.gridSize))
        raw = np.asarray(study.rawData, dtype=float)
        self.nCols = 1 if raw.ndim == 1 else int(raw.shape[1])
        if raw.ndim > 2:
            raise ConfigurationError('Data with more than two dimensions is not supported.')
        kind = getattr(om, 'deviceKind', KIND_TABLE)
        if type(om).pdf is not ObservationModel.pdf:  # a subclass that brings its own pdf wins over the op-code
            kind = KIND_TABLE
        if len(self.gridSize) > 2:
            raise NotImplementedError('The engine supports observation models with one or two parameters.')
        if kind == KIND_GAUSSIAN_MEAN and self.nCols != 2:
            raise ConfigurationError('GaussianMean expects data rows of the form [value, std].')
        if kind == KIND_POISSON and np.any(raw[np.isfinite(raw)] < 0):
            # observationModels.py:502 takes the factorial of the count: math.factorial raises for negative values
            raise ValueError('factorial() not defined for negative values')
        self.kind = kind
        self.plan = eng.plan(study.marginalGrid, study.latticeConstant, kind, om.segmentLength, self.nCols)
        self.data = eng.to_device(raw.reshape(len(raw), self.nCols), pinned=True)
        self.likTable = None
        if kind == KIND_TABLE:  # plugin without op-code: one host evaluation per time step, shared by all combos
            table = np.empty((self.T, self.G))
            for i in range(self.T):
                table[i] = np.asarray(om.processedPdf(study.grid, study.formattedData[i]), dtype=float).ravel()
            self.likTable = eng.to_device(table, pinned=True)
        self.prior = eng.to_device(np.asarray(study._computePrior(silent=True), dtype=float).reshape(-1))
        self.resetBase = None
        self._study = study

    def reset_base(self):
        """Observation-model prior on the grid, normalised to sum 1 (what ChangePoint / Independent restore:
        transitionModels.py:300-310, :350-359)."""
        if self.resetBase is None:
            study = self._study
            prior = study.observationModel.prior
            if callable(prior):
                values = np.asarray(prior(*study.grid), dtype=float) * np.ones(study.gridSize)
            elif isinstance(prior, np.ndarray):
                values = np.array(prior, dtype=float)
            else:
                values = np.ones(study.gridSize)
            values = values / np.sum(values)
            self.resetBase = self.eng.to_device(values.reshape(-1))
        return self.resetBase


class Study(object):
    """Fit with fixed hyper-parameter values (reference: core.py:39-486)."""

    def __init__(self, silent=False, engine=None):
        self.observationModel = None
        self.transitionModel = None
        self.gridSize = []
        self.boundaries = []
        self.marginalGrid = []
        self.grid = []
        self.latticeConstant = []
        self.rawData = np.array([])
        self.formattedData = np.array([])
        self.rawTimestamps = None
        self.formattedTimestamps = None
        self.posteriorSequence = []
        self.posteriorMeanValues = []
        self.logEvidence = 0
        self.localEvidence = []
        self.selectedHyperParameters = []
        self.fitWarningCounter = 0
        self._engineOverride = engine
        if not silent:
            print('+ Created new study.')

    # ------------------------------------------------------------------------- device-resident results (row f4)
    # After a fit the [T x G] posterior sequence stays in HBM; reading `posteriorSequence` downloads it (once) and
    # releases the device copy.  Marginal distributions and the time-averaged posterior (core.py:886, :915, :979-980)
    # are computed on the device while it is there: only [T x n_axis] numbers cross PCIe instead of T x G.
    @property
    def posteriorSequence(self):
        dev = self.__dict__.get('_postDev')
        if dev is not None:
            eng, plan, tensor, shape = dev
            self.__dict__['_postHost'] = eng.to_host(tensor).reshape(shape)
            self.__dict__['_postDev'] = None
        return self.__dict__.get('_postHost', [])

    @posteriorSequence.setter
    def posteriorSequence(self, value):
        self.__dict__['_postHost'] = value
        self.__dict__['_postDev'] = None

    def _setDeviceSequence(self, eng, plan, tensor, shape):
        self.__dict__['_postHost'] = None
        self.__dict__['_postDev'] = (eng, plan, tensor, list(shape))

    def _deviceMarginal(self, axis, row=None, average=False):
        """Marginal over the other grid axis, computed in HBM: all rows ([T, n_axis]), one row, or the time average."""
        eng, plan, tensor, shape = self.__dict__['_postDev']
        T, G = shape[0], int(np.prod(shape[1:]))
        seq = tensor.reshape(T, G)
        n = shape[1 + axis]
        if average:
            mean = eng.empty(G)
            eng.time_average(plan, seq, T, mean)
            seq, T = mean.reshape(1, G), 1
        elif row is not None:
            seq, T = seq[row:row + 1], 1
        out = eng.empty((T, n))
        eng.marginal(plan, seq, T, axis, out)
        host = eng.to_host(out)
        return host if (row is None and not average) else host[0]

    # ------------------------------------------------------------------------------------------ plumbing
    def _engine(self):
        return self._engineOverride if self._engineOverride is not None else _engine.default_engine()

    @property
    def log10Evidence(self):
        return self.logEvidence / np.log(10)

    def loadExampleData(self, silent=False):
        self.rawData = np.array([int(c) for c in COAL_MINING_DISASTERS])
        self.rawTimestamps = np.arange(1852, 1962)
        if not silent:
            print('+ Successfully imported example data.')

    def loadData(self, array, timestamps=None, silent=False):
        if isinstance(array, np.ndarray):
            self.rawData = array
        elif isinstance(array, list):
            if not silent:
                print('! WARNING: Data supplied as list, not as Numpy array. Converting list to Numpy array '
                      '(dtype=float).')
            self.rawData = np.array(array, dtype=float)
        else:
            raise ConfigurationError('Data type not supported. Please provide data as Numpy array.')
        self.rawTimestamps = np.arange(len(self.rawData))
        if timestamps is not None:
            if len(timestamps) == len(array):
                self.rawTimestamps = np.array(timestamps)
            elif not silent:
                print('! WARNING: Number of timestamps does not match number of data points. Omitting timestamps.')
        if not silent:
            print('+ Successfully imported array.')

    def load(self, array, timestamps=None, silent=False):
        self.loadData(array, timestamps=timestamps, silent=silent)

    def setObservationModel(self, L, silent=False):
        """Attach the likelihood plugin and build the parameter grid from its parameter values."""
        self.observationModel = L
        self.marginalGrid, self.gridSize, self.boundaries, self.latticeConstant = [], [], [], []
        for values, name in zip(L.parameterValues, L.parameterNames):
            if values is None:
                try:
                    values = L.estimateParameterValues(name, self.rawData)
                except Exception:
                    raise ConfigurationError('Could not estimate parameter values for "{}".'.format(name))
                print('+ Estimated parameter interval for "{}": [{}, {}] ({} values).'
                      .format(name, values[0], values[-1], len(values)))
            values = np.array(values, dtype=float)
            self.marginalGrid.append(values)
            self.gridSize.append(len(values))
            self.boundaries.append([values[0], values[-1]])
            if is_regular(values):
                self.latticeConstant.append(np.abs(values[0] - values[1]))
            else:
                print('! WARNING: Supplied parameter values for "{}" are not equally spaced. Assuming categorical '
                      'parameter.'.format(name))
                self.latticeConstant.append(1.)
        self.grid = list(np.meshgrid(*self.marginalGrid, indexing='ij'))
        if self.transitionModel is not None:
            self.transitionModel.latticeConstant = self.latticeConstant
        if not silent:
            print('+ Observation model: {}. Parameter(s): {}'.format(L, L.parameterNames))

    def setOM(self, L, silent=False):
        self.setObservationModel(L, silent=silent)

    def setTransitionModel(self, T, silent=False):
        if str(T) == 'Break-point':
            raise ConfigurationError('The "BreakPoint" transition model can only be used with the '
                                     '"SerialTransitionModel" class.')
        self.transitionModel = T
        T.study = weakref.proxy(self)  # back-pointer of core.py:281 without a reference cycle: results are freed on release
        T.latticeConstant = self.latticeConstant
        if not silent:
            print('+ Transition model: {}. Hyper-Parameter(s): {}'
                  .format(T, self._unpackAllHyperParameters(values=False)))

    def setTM(self, T, silent=False):
        self.setTransitionModel(T, silent=silent)

    def __getstate__(self):
        self.posteriorSequence  # noqa: B018 -- brings a device-resident sequence to the host
        d = self.__dict__.copy()
        d['_engineOverride'] = None  # engines hold library handles and a device: not part of a saved study
        d['_postDev'] = None
        return d

    def _attachBackPointers(self):
        """(Re-)attach the weak `study` back-pointer of every transition model of this study (dropped by pickling)."""
        def walk(model):
            if model is None:
                return
            model.study = weakref.proxy(self)
            for sub in getattr(model, 'models', []):
                walk(sub)
        walk(self.transitionModel)
        for model in getattr(self, 'transitionModels', []):
            walk(model)

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._attachBackPointers()

    def set(self, *args, **kwargs):
        for key in kwargs:
            if key != 'silent':
                raise TypeError("set() got an unexpected keyword argument '{}'".format(key))
        silent = kwargs.get('silent', False)
        seen = set()
        for model in args:
            if isinstance(model, ObservationModel):
                role, setter = 'observation', self.setObservationModel
            elif isinstance(model, TransitionModel):
                role, setter = 'transition', self.setTransitionModel
            else:
                raise ConfigurationError('Expected observation model or transition model instance as first argument.')
            if role in seen:
                raise ConfigurationError('More than one {} model supplied.'.format(role))
            seen.add(role)
            setter(model, silent=silent)

    def _computePrior(self, silent=False):
        """Initial parameter distribution on the grid with the reference's normalisation conventions
        (core.py:184-265): flat / array / callable priors end up summing to 1/prod(latticeConstant) -- arrays are
        normalised IN PLACE like the reference does -- SymPy random variables are taken as densities as they are."""
        prior = self.observationModel.prior
        cell = np.prod(self.latticeConstant)
        if prior is None:
            flat = np.ones(self.gridSize)
            return flat / np.sum(flat) / cell
        if isinstance(prior, np.ndarray):
            if prior.shape != tuple(self.gridSize):
                raise ConfigurationError('Prior array does not match parameter grid size.')
            total = np.sum(prior)
            if total != 1.:
                prior /= total
                prior /= cell
            return prior
        if callable(prior):
            values = prior(*self.grid) * np.ones(self.gridSize)
            total = np.sum(values)
            if total != 1.:
                values = values / total / cell
            return values
        rvs = [prior] if _is_sympy_rv(prior) else prior
        if isinstance(rvs, (list, tuple)):
            import string

            import sympy
            if len(rvs) != len(self.observationModel.parameterNames):
                raise ConfigurationError('Observation model contains {} parameters, but {} priors were provided.'
                                         .format(len(self.observationModel.parameterNames), len(rvs)))
            values = 1.
            for letter, rv, axisGrid in zip(string.ascii_lowercase, rvs, self.grid):
                if not _is_sympy_rv(rv):
                    raise ConfigurationError('Only lambda functions or SymPy random variables can be used as a prior.')
                if len(_free_symbols(rv)) > 0:
                    raise ConfigurationError('Prior distribution must not contain free parameters.')
                values = values * _sympy_density(rv, sympy.Symbol(letter))(axisGrid)
            return np.asarray(values, dtype=float) * np.ones(self.gridSize)
        raise ConfigurationError('Unsupported prior specification.')

    # ------------------------------------------------------------------------ hyper-parameter tree plumbing
    def _unpackHyperParameters(self, transitionModel, values=False):
        out = [self._unpackHyperParameters(m, values=values) for m in getattr(transitionModel, 'models', [])]
        if hasattr(transitionModel, 'hyperParameterNames'):
            out.extend(transitionModel.hyperParameterValues if values else transitionModel.hyperParameterNames)
        return out

    def _unpackAllHyperParameters(self, values=True):
        return list(flatten(self._unpackHyperParameters(self.transitionModel, values=values)))

    def _locate(self, nameTree, name):
        path = recursiveIndex(nameTree, name)
        if len(path) == 0:
            raise ConfigurationError('Could not find any hyper-parameter named {}.'.format(name))
        model = self.transitionModel
        for pos in path[:-1]:
            model = model.models[pos]
        assignNestedItem(nameTree, path, ' ')  # a name listed twice addresses consecutive occurrences
        return model, model.hyperParameterNames.index(name)

    def _unpackSelectedHyperParameters(self):
        if not self.selectedHyperParameters:
            return self._unpackAllHyperParameters()
        tree = self._unpackHyperParameters(self.transitionModel)
        out = []
        for name in self.selectedHyperParameters:
            model, pos = self._locate(tree, name)
            out.append(model.hyperParameterValues[pos])
        return out

    def _setAllHyperParameters(self, x):
        tree = self._unpackHyperParameters(self.transitionModel)
        names = list(flatten(self._unpackHyperParameters(self.transitionModel)))
        for name, value in zip(names, list(x)):
            model, pos = self._locate(tree, name)
            model.hyperParameterValues[pos] = value

    def _setSelectedHyperParameters(self, x):
        if not self.selectedHyperParameters:
            self._setAllHyperParameters(x)
            return 1
        tree = self._unpackHyperParameters(self.transitionModel)
        for name, value in zip(self.selectedHyperParameters, list(x)):
            model, pos = self._locate(tree, name)
            model.hyperParameterValues[pos] = value
        return 1

    def _unpackLabelled(self, transitionModel, label):
        out = [self._unpackLabelled(m, label) for m in getattr(transitionModel, 'models', [])]
        if hasattr(transitionModel, 'hyperParameterNames') and str(transitionModel) == label:
            out.extend(transitionModel.hyperParameterNames)
        return out

    def _unpackChangepointNames(self, transitionModel):
        return self._unpackLabelled(transitionModel, 'Change-point')

    def _unpackBreakpointNames(self, transitionModel):
        return self._unpackLabelled(transitionModel, 'Serial transition model')

    def _getHyperParameterIndex(self, transitionModel, name):
        names = list(flatten(self._unpackHyperParameters(transitionModel, values=False)))
        if name not in names:
            raise PostProcessingError('Could not find any hyper-parameter with name: {}.'.format(name))
        return names.index(name)

    def getHyperParameterValue(self, name):
        return self._unpackAllHyperParameters(values=True)[self._getHyperParameterIndex(self.transitionModel, name)]

    def _checkConsistency(self):
        if len(self.rawData) == 0:
            raise ConfigurationError('No data loaded.')
        if not self.observationModel:
            raise ConfigurationError('No observation model chosen.')
        if not self.transitionModel:
            raise ConfigurationError('No transition model chosen.')
        names = self._unpackAllHyperParameters(values=False)
        twice = sorted({n for n in names if names.count(n) > 1})
        if twice:
            raise ConfigurationError('Detected duplicate hyper-parameter names: {}.'.format(twice))

    def _formatData(self):
        seg = self.observationModel.segmentLength
        self.formattedData = movingWindow(self.rawData, seg)
        self.formattedTimestamps = self.rawTimestamps[seg - 1:]

    # ---------------------------------------------------------------------------------------- device sweep
    def _lower(self, hyperRows, timestamps, online=False, model=None):
        model = self.transitionModel if model is None else model
        _tm.assign_columns(model)
        ctx = _tm.LoweringContext(self.observationModel.parameterNames, self.latticeConstant, hyperRows, timestamps,
                                  online=online)
        model.lower(ctx, _tm.Window.everything(ctx.B))
        return ctx

    def _warnZeroNorm(self, backward):
        if self.fitWarningCounter < 5:
            print('    ! WARNING: {} distribution contains only zeros, check parameter boundaries!'
                  .format('Posterior' if backward else 'Forward pass'))
            print('      Stopping inference process. Setting model evidence to zero.')
        elif self.fitWarningCounter == 5:
            print('    ! WARNING: Will omit further warnings about parameter boundaries.')
        self.fitWarningCounter += 1

    def _fitSingle(self, forwardOnly, evidenceOnly, silent):
        """One combination of hyper-parameter values: forward filter, backward smoother, means (core.py:330-486)."""
        eng = self._engine()
        if len(self.formattedData) == 0:
            # fewer data points than one segment: the loops of core.py:372-470 do not run, the evidence is the
            # constant of core.py:417 and every sequence is empty
            self.logEvidence = float(np.log(np.prod(self.latticeConstant)))
            self.localEvidence = np.empty(0)
            self.posteriorSequence = np.empty([0] + list(self.gridSize))
            self.posteriorMeanValues = np.empty([len(self.gridSize), 0])
            return
        ses = _Session(self, eng)
        T, G = ses.T, ses.G
        values = self._unpackAllHyperParameters()
        for v in values:
            if isinstance(v, (Iterable, str)) and not np.isscalar(v):
                raise ConfigurationError('Study.fit needs scalar hyper-parameter values; use HyperStudy for lists.')
        ctx = self._lower(np.array([values], dtype=float).reshape(1, len(values)), self.formattedTimestamps)
        program = _engine.Program(eng, ctx.ops, 1)
        logE, local = eng.zeros(1), eng.zeros((1, T))
        alive = eng.zeros(1, dtype=torch.int32)
        seq = None if evidenceOnly else eng.empty((1, T, G))
        common = dict(T=T, B=1, data=ses.data, prior=ses.prior, lik_table=ses.likTable, program=program,
                      reset_base=ses.reset_base() if ctx.usesReset else None, log_evidence=logE,
                      local_evidence=local, alive=alive, alpha_seq=seq)
        # a backward pass follows: the filtering rows may stay unnormalised (the smoother is scale-free per row)
        fwdFlags = _engine.F_EVIDENCE_ONLY if evidenceOnly else (0 if forwardOnly else _engine.F_RAW_ALPHA)
        eng.run('forward', ses.plan, fwdFlags, **common)
        state = int(eng.to_host(alive)[0])
        rawRows = False
        if state == 1 and not (forwardOnly or evidenceOnly):
            # smoothed rows may come back unnormalised (+ their scale): finalize normalises them below
            eng.run('backward', ses.plan, _engine.F_RAW_POSTERIOR, row_scale=eng.empty((1, T)), **common)
            rawRows = True
            state = int(eng.to_host(alive)[0])
        self.localEvidence = eng.to_host(local)[0]
        self.logEvidence = float(eng.to_host(logE)[0])
        if not silent and state != 0:
            print('    + Finished forward pass.')
            print('    + Log10-evidence: {:.5f}'.format(self.logEvidence / np.log(10)))
        if state != 1:
            self._warnZeroNorm(backward=(state == -1))
            self.logEvidence = -np.inf
            if not evidenceOnly:
                # the reference leaves np.empty memory behind the step that died (core.py:399-400); what was computed up
                # to there is exposed as distributions: rows stored raw (F_RAW_ALPHA / F_RAW_POSTERIOR) are normalised
                if not forwardOnly:
                    eng.finalize(ses.plan, seq, T, eng.empty((len(self.gridSize), T)), _engine.F_NORMALIZE_ROWS)
                self.posteriorSequence = eng.to_host(seq).reshape([T] + self.gridSize)
            return
        if evidenceOnly:
            self.posteriorMeanValues = []
            return
        means = eng.empty((len(self.gridSize), T))
        eng.finalize(ses.plan, seq, T, means, _engine.F_NORMALIZE_ROWS if rawRows else 0)
        self._setDeviceSequence(eng, ses.plan, seq[0], [T] + list(self.gridSize))
        self.posteriorMeanValues = eng.to_host(means)
        if not silent:
            if not forwardOnly:
                print('    + Finished backward pass.')
            print('    + Computed mean parameter values.')

    def fit(self, forwardOnly=False, evidenceOnly=False, silent=False):
        """Sequence of posterior distributions and the model evidence (same contract as core.py:330-343)."""
        self._checkConsistency()
        if not silent:
            print('+ Started new fit:')
        self._formatData()
        if not silent:
            print('    + Formatted data.')
        self.logEvidence = 0
        self._fitSingle(forwardOnly, evidenceOnly, silent)

    # -------------------------------------------------------------------------------------------- optimize
    def optimize(self, parameterList=[], forwardOnly=False, **kwargs):
        """Maximise the log-evidence over hyper-parameters with SciPy's COBYLA (core.py:488-565); every
        evaluation is one evidence-only device pass."""
        from scipy.optimize import minimize
        self.selectedHyperParameters = [parameterList] if isinstance(parameterList, str) else parameterList
        print('+ Starting optimization...')
        self._checkConsistency()
        if self.selectedHyperParameters:
            print('  --> Parameter(s) to optimize:', self.selectedHyperParameters)
        else:
            print('  --> All model parameters are optimized (except change/break-points).')
            everything = list(flatten(self._unpackHyperParameters(self.transitionModel)))
            points = list(flatten(self._unpackChangepointNames(self.transitionModel))) + \
                list(flatten(self._unpackBreakpointNames(self.transitionModel)))
            self.selectedHyperParameters = [x for x in everything if x not in points]
        x0 = self._unpackSelectedHyperParameters()
        if len(x0) == 0:
            self.selectedHyperParameters = []
            raise ConfigurationError('No parameters to optimize. Check parameter names.')
        result = minimize(self._optimizationStep, x0, method='COBYLA', **kwargs)
        print('+ Finished optimization.')
        self._setSelectedHyperParameters(result.x)
        self.fit(forwardOnly=forwardOnly)
        self.selectedHyperParameters = []

    def _optimizationStep(self, x):
        self._setSelectedHyperParameters(x)
        self.fit(evidenceOnly=True, silent=True)
        print('    + Log10-evidence: {:.5f}'.format(self.logEvidence / np.log(10)), '- Parameter values:', x)
        return -self.logEvidence

    # -------------------------------------------------------------------------------------------- simulate
    def simulate(self, x, t=None, density=False):
        """Probability (density) of observation values `x` under the posterior of time stamp `t` (or the time-averaged
        posterior for t=None), reference: core.py:567-602 -- prob[i] = sum_g pdf(x[i] | g) * post[g].

        On the device this is ONE evidence-only forward pass over the pseudo series `x` whose transition program
        restores `post` before every step (the RESET operator of the Independent model, transitionModels.py:339-360):
        the evidence increment of step i is exactly sum_g pdf(x[i] | g) * post[g], so no extra kernel is needed."""
        om = self.observationModel
        if om.segmentLength > 1:
            raise NotImplementedError('Method "simulate" is only available for observation models with '
                                      'segment length 1.')
        seq = np.asarray(self.posteriorSequence, dtype=float)
        if seq.ndim < 2 or seq.size == 0:
            raise PostProcessingError('Cannot simulate observations as the posterior sequence has not yet been '
                                      'computed. Run complete fit.')
        if t is None:
            post = np.sum(seq, axis=0) / len(seq)
        else:
            stamps = list(self.formattedTimestamps)
            if t not in stamps:
                raise PostProcessingError('Supplied time ({}) does not exist in data or is out of range.'.format(t))
            post = seq[stamps.index(t)]
        eng = self._engine()
        x = np.asarray(x, dtype=float)
        n = len(x)
        G = int(np.prod(self.gridSize))
        nCols = 1 if x.ndim == 1 else int(x.shape[1])
        kind = getattr(om, 'deviceKind', KIND_TABLE)
        if type(om).pdf is not ObservationModel.pdf:
            kind = KIND_TABLE
        plan = eng.plan(self.marginalGrid, self.latticeConstant, kind, 1, nCols)
        table = None
        if kind == KIND_TABLE:
            table = np.array([np.asarray(om.pdf(self.grid, [xi]), dtype=float).ravel() for xi in x])
        base = eng.to_device(post.reshape(-1))
        always = np.array([[_tm.ALWAYS[0], _tm.ALWAYS[1], _tm.ALWAYS[0], _tm.ALWAYS[1]]], dtype=np.int32)
        program = _engine.Program(eng, [dict(kind=_tm.OP_RESET, axis=0, param=np.ones(1),
                                             radius=np.zeros(1, dtype=np.int32), window=always)], 1)
        prob = np.zeros(n)
        lo = 0
        while lo < n:  # one pass; a value with zero probability ends a pass (core.py:388-400) and starts the next one
            m = n - lo
            local, logE = eng.zeros((1, m)), eng.zeros(1)
            alive = eng.zeros(1, dtype=torch.int32)
            eng.run('forward', plan, _engine.F_EVIDENCE_ONLY, T=m, B=1,
                    data=eng.to_device(x[lo:].reshape(m, nCols)), prior=base, reset_base=base,
                    lik_table=None if table is None else eng.to_device(table[lo:]), program=program,
                    log_evidence=logE, local_evidence=local, alive=alive)
            got = eng.to_host(local)[0] / np.prod(self.latticeConstant)
            if int(eng.to_host(alive)[0]) == 1:
                prob[lo:] = got
                break
            dead = int(np.flatnonzero(~(got > 0))[0])
            prob[lo:lo + dead] = got[:dead]
            lo += dead + 1
        if not density:
            prob /= np.sum(prob)
        return prob

    # ------------------------------------------------------------------------------------------- accessors
    def _parameterIndex(self, name):
        names = list(self.observationModel.parameterNames)
        if name not in names:
            raise PostProcessingError('Wrong parameter name. Available options: {0}'.format(names))
        return names.index(name)

    def _hasPosterior(self):
        if self.__dict__.get('_postDev') is not None:
            return True
        return isinstance(self.posteriorSequence, np.ndarray) and self.posteriorSequence.size > 0

    def getParameterMeanValues(self, name):
        return self.posteriorMeanValues[self._parameterIndex(name)]

    def getParameterDistribution(self, t, name, plot=False, density=True, **kwargs):
        """Marginal distribution of one parameter at time stamp `t` (or 'avg')."""
        if not self._hasPosterior():
            raise PostProcessingError('Cannot plot posterior sequence as it has not yet been computed. '
                                      'Run complete fit.')
        axis = self._parameterIndex(name)
        average = isinstance(t, str) and t == 'avg'
        if not average:
            stamps = list(self.formattedTimestamps)
            if t not in stamps:
                raise PostProcessingError('Supplied time ({}) does not exist in data or is out of range.'.format(t))
        if self.__dict__.get('_postDev') is not None:  # sequence still in HBM: reduce it there
            marginal = self._deviceMarginal(axis, row=None if average else stamps.index(t), average=average)
        else:
            if average:
                dist = np.sum(self.posteriorSequence, axis=0) / len(self.posteriorSequence)
            else:
                dist = self.posteriorSequence[stamps.index(t)]
            others = tuple(a for a in range(dist.ndim) if a != axis)
            marginal = np.sum(dist, axis=others) if others else dist.copy()
        if density:
            marginal = marginal / self.latticeConstant[axis]
        return self.marginalGrid[axis], marginal

    def getPD(self, t, name, plot=False, density=True, **kwargs):
        return self.getParameterDistribution(t, name, plot=plot, density=density, **kwargs)

    def getParameterDistributions(self, name, plot=False, density=True, **kwargs):
        """Time series of marginal distributions of one parameter: array [T, n_axis]."""
        if not self._hasPosterior():
            raise PostProcessingError('Cannot plot posterior sequence as it has not yet been computed. '
                                      'Run complete fit.')
        axis = self._parameterIndex(name)
        if self.__dict__.get('_postDev') is not None:  # sequence still in HBM: only [T x n_axis] numbers come back
            marginal = self._deviceMarginal(axis)
        else:
            seq = np.asarray(self.posteriorSequence)
            others = tuple(a + 1 for a in range(seq.ndim - 1) if a != axis)
            marginal = np.sum(seq, axis=others) if others else seq.copy()
        if density:
            marginal = marginal / self.latticeConstant[axis]
        return self.marginalGrid[axis], marginal

    def getPDs(self, name, plot=False, density=True, **kwargs):
        return self.getParameterDistributions(name, plot=plot, density=density, **kwargs)


class HyperStudy(Study):
    """Sweep over a grid of hyper-parameter values with evidence-weighted model averaging
    (reference: core.py:1118-1495).  The sweep is the batch axis of the device kernels and, under
    torch.distributed, is sharded across ranks/GPUs (bayesloop_b200/distributed.py)."""

    shareDeal = None  # under torch.distributed: 'groups' / 'changepoints' instead of the choice of _deal_shared

    def __init__(self, silent=False, engine=None):
        super(HyperStudy, self).__init__(silent=silent, engine=engine)
        self.hyperGrid = []
        self.hyperGridValues = []
        self.hyperGridConstant = []
        self.flatHyperParameters = []
        self.flatHyperParameterNames = []
        self.flatHyperPriors = []
        self.flatHyperPriorValues = []
        self.hyperParameterDistribution = None
        self.averagePosteriorSequence = None
        self.logEvidenceList = []
        self.localEvidenceList = []
        self.sweepStats = {}
        self.maxWave = None
        self.shareChangepoints = True  # change-point prefix sharing where the sweep allows it (_share_structure)
        if not silent:
            print('  --> Hyper-study')

    def _unpackHyperPriors(self, transitionModel):
        out = [self._unpackHyperPriors(m) for m in getattr(transitionModel, 'models', [])]
        if hasattr(transitionModel, 'prior'):
            if len(getattr(transitionModel, 'hyperParameterNames', [])) > 0 or str(transitionModel) == 'Break-point':
                out.append(transitionModel.prior)
        return out

    def _unpackAllHyperPriors(self):
        return list(flatten(self._unpackHyperPriors(self.transitionModel)))

    def _createHyperGrid(self, silent=False):
        """Cartesian grid of hyper-parameter values, its lattice constants and the joint hyper-prior
        (semantics of core.py:1142-1245)."""
        self.flatHyperParameters = self._unpackAllHyperParameters()
        self.flatHyperParameterNames = self._unpackAllHyperParameters(values=False)
        self.flatHyperPriors = self._unpackAllHyperPriors()
        for i, v in enumerate(self.flatHyperParameters):
            if isinstance(v, str) and v == 'all':
                self.flatHyperParameters[i] = self.formattedTimestamps[:-1]

        if len(self.flatHyperParameterNames) > 0:
            mesh = np.meshgrid(*self.flatHyperParameters, indexing='ij')
            self.hyperGridValues = np.array([m.ravel() for m in mesh]).T
        else:
            self.hyperGridValues = np.array([])

        constants = []
        for values in self.flatHyperParameters:
            step = 1
            if isinstance(values, Iterable) and len(values) > 1:
                arr = np.array(values)
                if is_regular(arr, tol=1e-10) and not np.any(np.abs(np.diff(arr, 2)) >= 1e-10):
                    step = np.abs(arr[1] - arr[0])
            constants.append(step)
        self.hyperGridConstant = np.array(constants)

        perAxis, labels = [], []
        for prior, values, const, name in zip(self.flatHyperPriors, self.flatHyperParameters,
                                              self.hyperGridConstant, self.flatHyperParameterNames):
            if prior is None:
                p = np.ones_like(values, dtype=float)
                p = p / np.sum(p) / const
                labels.append('uniform')
            elif callable(prior):
                try:
                    p = np.array([prior(v) for v in values])
                    total = np.sum(p)
                    p = p / total / const
                except Exception:
                    raise ConfigurationError('Failed to set hyper-prior for "{}" from function "{}".'
                                             .format(name, getattr(prior, '__name__', prior)))
                labels.append(prior.__name__ + (' (re-normalized)' if total != 1. else ''))
            elif isinstance(prior, Iterable):
                if len(prior) != len(values):
                    raise ConfigurationError('Failed to set hyper-prior for "{}" from list/array.'.format(name))
                total = np.sum(prior)
                if isinstance(prior, np.ndarray) and prior.dtype.kind == 'f':
                    prior /= total  # in place, like the reference (core.py:1210-1213)
                    prior /= const
                    p = prior
                else:
                    p = np.array(prior, dtype=float) / total / const
                labels.append('list/array' + (' (re-normalized)' if total != 1. else ''))
            else:  # SymPy random variable: density evaluated as is
                import sympy.abc
                if len(_free_symbols(prior)) > 0:
                    raise ConfigurationError('Hyper-prior for "{}" must not contain free parameters.'.format(name))
                p = _sympy_density(prior, sympy.abc.x)(values)
                labels.append('sympy')
            perAxis.append(p)

        if len(self.flatHyperParameterNames) > 0:
            mesh = np.meshgrid(*perAxis, indexing='ij')
            self.flatHyperPriorValues = np.prod(np.array([m.ravel() for m in mesh]).T, axis=1)
            if not silent and len(self.hyperGridValues) > 1:
                print('+ Set hyper-prior(s): {}'.format(labels))
        else:
            self.flatHyperPriorValues = np.array([1])

    def _prepareSweep(self, forwardOnly, evidenceOnly):
        """Device-resident inputs and work buffers of a sweep over this rank's rows of self.hyperGridValues:
        session (data, prior, tables), lowered transition program, wave buffer sized to the free HBM."""
        from . import distributed as dist
        eng = self._engine()
        ses = _Session(self, eng)
        T, G = ses.T, ses.G
        Ball = len(self.hyperGridValues)
        hyperAll = np.asarray(self.hyperGridValues, dtype=float).reshape(Ball, -1)
        ctx = self._lower(hyperAll, self.formattedTimestamps)
        share = None
        if self.shareChangepoints and not (forwardOnly or evidenceOnly):
            share = _share_structure(ctx.ops, T)
        if share is None:
            rows = dist.deal_by_cost(_combo_cost(ctx.ops, Ball, T))
        else:
            # whole groups or whole change-points dealt round-robin over the ranks (_deal_shared picks the cheaper deal)
            rank, size = dist.world()
            deal = 'groups' if size == 1 else self.shareDeal
            if deal is None:
                # the decision must be the same on every rank: smallest free memory of the ranks, grid-level slot count
                free = dist.min_over_ranks(eng, eng.free_bytes())
                fits = 2 * share['nG'] * T * G * 8 <= (int(free * 0.85) - T * G * 8) // 2
                slots = 15 if G * 8 > 200 * 1024 else 4 * max(eng.sm_count(), 148)  # clusters of 8 SMs / chains per SM
                deal = _deal_shared(share, ctx.ops, size, slots, fits)
            elif (deal == 'groups' and share['nG'] < size) or (deal == 'changepoints' and share['nC'] // size < 2):
                deal = None
            if deal is None:
                share, rows = None, dist.deal_by_cost(_combo_cost(ctx.ops, Ball, T))
            else:
                # this rank's combinations are laid out change-point-major: slot = (change-point index) * nG + (group index)
                gsel = np.arange(rank, share['nG'], size) if deal == 'groups' else np.arange(share['nG'])
                csel = np.arange(share['nC']) if deal == 'groups' else np.arange(rank, share['nC'], size)
                where = np.full((share['nC'], share['nG']), -1, dtype=np.int64)
                where[share['cidx'], share['group']] = np.arange(Ball)
                rows = where[np.ix_(csel, gsel)].reshape(-1)
                share = dict(share, nG=len(gsel), nC=len(csel), cvals=share['cvals'][csel], deal=deal)
        B = len(rows)
        hp = np.asarray(self.flatHyperPriorValues, dtype=float)[rows]
        ops = [dict(op, param=np.asarray(op['param'])[rows], radius=np.asarray(op['radius'])[rows],
                    window=np.asarray(op['window'])[rows]) for op in ctx.ops]
        ctx.ops = ops
        sw = dict(eng=eng, ses=ses, T=T, G=G, Ball=Ball, rows=rows, B=B, hp=hp, ops=ctx.ops,
                  forwardOnly=forwardOnly, evidenceOnly=evidenceOnly, share=share)
        sw['program'] = _engine.Program(eng, ctx.ops, B)
        sw['resetBase'] = ses.reset_base() if ctx.usesReset else None
        sw['logE'], sw['local'] = eng.zeros(max(B, 1)), eng.zeros((max(B, 1), T))
        sw['alive'] = eng.zeros(max(B, 1), dtype=torch.int32)
        sw['hpDev'] = eng.to_device(hp) if B > 0 else None
        sw['part'] = eng.zeros(T)
        if evidenceOnly:
            sw['wave'], sw['buf'], sw['avg'], sw['means'], sw['rowScale'] = max(B, 1), None, None, None, None
        else:
            sw['avg'] = eng.zeros((T, G))
            sw['means'] = eng.empty((len(self.gridSize), T))
            budget = int(eng.free_bytes() * 0.85) - T * G * 8
            if share is not None:
                self._prepareShared(sw, budget)
                budget -= 2 * share['nG'] * T * G * 8
            sw['wave'] = int(max(1, min(B, budget // max(1, T * G * 8))))
            if share is not None:  # whole change-points per wave
                sw['wave'] = max(1, sw['wave'] // share['nG']) * share['nG']
            if self.maxWave:  # user cap on the combinations fitted concurrently (memory, or to exercise the wave logic)
                sw['wave'] = int(max(1, min(sw['wave'], self.maxWave)))
                if share is not None:
                    sw['wave'] = max(1, sw['wave'] // share['nG']) * share['nG']
            sw['buf'] = eng.empty((sw['wave'], T, G)) if B > 0 else None
            sw['rowScale'] = eng.empty((sw['wave'], T)) if (B > 0 and not forwardOnly) else None
        with np.errstate(divide='ignore'):
            sw['logHp'] = np.log(hp)
        sw['logHpDev'] = eng.to_device(sw['logHp']) if B > 0 else None
        sw['shift'] = eng.full(1, -math.inf)          # running reference log-weight of the average (device scalar)
        sw['weights'] = eng.zeros(max(B, 1))          # log-weights of the combos relative to it
        return sw

    def _prepareShared(self, sw, budget):
        """Programs and buffers of the change-point prefix sharing (see _share_structure / _executeSharedSweep)."""
        eng, T, G, share = sw['eng'], sw['T'], sw['G'], sw['share']
        nG, k = share['nG'], share['k']
        if 2 * nG * T * G * 8 > budget // 2:  # the two shared sequences must leave room for the combinations
            sw['share'] = None
            return
        first = np.arange(nG)  # slots of the first change-point: one representative per group

        def variant(window):
            out = []
            for j, op in enumerate(sw['ops']):
                w = np.asarray(op['window'])[first].copy()
                if j == k:
                    w[:] = window
                out.append(dict(op, param=np.asarray(op['param'])[first], radius=np.asarray(op['radius'])[first], window=w))
            return _engine.Program(eng, out, nG)

        sw['progShared'] = variant([0, 0, 0, 0])  # the reset never fires: the change-point-free model of each group
        sw['progSuffix'] = variant([0, 1, 1, 2])  # reset after the FIRST step of a window that starts at the change-point
        sw['alphaS'], sw['ratio'] = eng.empty((nG, T, G)), eng.empty((nG, T, G))
        sw['localS'], sw['scratchLocal'] = eng.zeros((nG, T)), eng.zeros((nG, T))
        sw['rowScaleS'], sw['logES'] = eng.empty((nG, T)), eng.zeros(nG)
        sw['aliveS'] = eng.zeros(nG, dtype=torch.int32)
        sw['saveRow'], sw['scratchLogE'] = eng.empty((nG, G)), eng.zeros(nG)
        sw['cOfSlot'] = eng.to_device(np.repeat(share['cvals'], nG))
        sw['cpDev'] = eng.to_device(np.asarray(share['cvals'], dtype=np.int32))
        sw['gOfSlot'] = eng.to_device(np.tile(np.arange(nG), share['nC']))

    def _executeSharedSweep(self, sw):
        """Sweep with change-point prefix sharing.  Combination (g, c) = group g with its
        reset after step c.  Because the reset erases the history (transitionModels.py:300-312):
          * filtering rows t <= c and the evidence increments up to c are those of the group's change-point-free run;
          * filtering rows t > c come from a forward pass over the steps c .. T-1 only (its first row is discarded);
          * the backward message of rows t > c is the change-point-free run's, so smoothed row = alpha(g, c)[t] * beta(g)[t]
            (blg_share_apply; `ratio` below holds beta up to a factor per row);
          * smoothed rows t <= c come from a backward pass over the rows 0 .. c+1 only, reading the group's shared
            filtering rows out of place (alpha_src); the reset fires when row c+1 is processed, so what that row holds is
            irrelevant; the combination's own row c+1 is restored after.
        Executed cell updates per combination: (T - c) + (c + 2) instead of 2 T.  All passes are the ordinary kernels
        on WINDOWS of the sequences (seq_stride / row_stride of include/blgrid.h)."""
        from . import distributed as dist
        eng, ses, T, G, B = sw['eng'], sw['ses'], sw['T'], sw['G'], sw['B']
        share = sw['share']
        nG, nC, cvals = share['nG'], share['nC'], [int(c) for c in share['cvals']]
        plan, buf, avg, local, alive, logE = ses.plan, sw['buf'], sw['avg'], sw['local'], sw['alive'], sw['logE']
        rowScale, shift, weights = sw['rowScale'], sw['shift'], sw['weights']
        alphaS, ratio, localS = sw['alphaS'], sw['ratio'], sw['localS']
        lc = float(np.prod(self.latticeConstant))
        avg.zero_()
        shift.fill_(-math.inf)
        base = dict(prior=ses.prior, reset_base=sw['resetBase'])
        executed = 0

        def window(t0):  # inputs of a call whose first row is time step t0
            return dict(data=ses.data[t0:], lik_table=None if ses.likTable is None else ses.likTable[t0:])

        # 1. the change-point-free run of every group: filtering rows, and the backward message itself -- the smoother
        #    run on a sequence of ONES returns alpha * beta = beta row by row (the beta recursion, core.py:467-470, does
        #    not depend on alpha), exact where a quotient posterior / alpha would lose the cells whose alpha underflows
        shared = dict(T=T, B=nG, program=sw['progShared'], log_evidence=sw['logES'], alive=sw['aliveS'], **base, **window(0))
        eng.run('forward', plan, _engine.F_RAW_ALPHA, alpha_seq=alphaS, local_evidence=localS, **shared)
        ratio.fill_(1.0)
        eng.run('backward', plan, _engine.F_RAW_POSTERIOR, alpha_seq=ratio, local_evidence=sw['scratchLocal'],
                row_scale=sw['rowScaleS'], **shared)
        executed += 2 * nG * T
        if not bool((sw['aliveS'] == 1).all()):
            return None  # a group whose change-point-free run dies: let the plain sweep sort out who survives
        prefix = torch.cumsum(torch.log(localS / lc), dim=1)  # [nG][T]: log-evidence of the steps 0 .. t

        perWave = sw['wave'] // nG  # change-points per wave
        steps = torch.arange(T, device=local.device)
        waves = 0
        for ci0 in range(0, nC, perWave):
            ci1 = min(nC, ci0 + perWave)
            s0, nb = ci0 * nG, (ci1 - ci0) * nG
            waves += 1
            # 2a. filtering rows after the change-point: forward pass over the steps c .. T-1
            for ci in range(ci0, ci1):
                c, j = cvals[ci], (ci - ci0) * nG
                eng.run('forward', plan, _engine.F_RAW_ALPHA, T=T - c, B=nG, program=sw['progSuffix'],
                        alpha_seq=buf[j:j + nG, c:], seq_stride=T * G, local_evidence=local[s0 + j:s0 + j + nG, c:],
                        row_stride=T, log_evidence=sw['scratchLogE'], alive=alive[s0 + j:s0 + j + nG], **base, **window(c))
                executed += nG * (T - c)
            # evidence of every combination: shared steps 0 .. c, own steps c+1 .. T-1 (core.py:403, :417)
            cs, gs = sw['cOfSlot'][s0:s0 + nb], sw['gOfSlot'][s0:s0 + nb]
            own = torch.where(steps[None, :] > cs[:, None], torch.log(local[s0:s0 + nb] / lc), torch.zeros((), device=local.device,
                                                                                                       dtype=local.dtype))
            le = prefix[gs, cs] + own.sum(dim=1) + math.log(lc)
            logE[s0:s0 + nb] = torch.where(alive[s0:s0 + nb] == 1, le, torch.full_like(le, -math.inf))
            eng.wave_weights(plan, logE[s0:s0 + nb], sw['logHpDev'][s0:s0 + nb], nb, shift, avg, T * G, weights[s0:s0 + nb])
            # 2b. smoothed rows after the change-point, all change-points of the wave in one call: own filtering rows x
            #     the group's backward message (read once per group and row)
            eng.share_apply(plan, ratio, T * G, nG, sw['cpDev'][ci0:ci1], T=T, B=nb, alpha_seq=buf, seq_stride=T * G,
                            row_scale=rowScale, local_evidence=local[s0:s0 + nb], row_stride=T, alive=alive[s0:s0 + nb],
                            log_evidence=logE[s0:s0 + nb], program=sw['progSuffix'], **base, **window(0))
            #     ... and before it: backward pass over the rows 0 .. c+1
            for ci in range(ci0, ci1):
                c, j = cvals[ci], (ci - ci0) * nG
                sl = slice(s0 + j, s0 + j + nG)
                sw['saveRow'].copy_(buf[j:j + nG, c + 1])
                keepScale, keepLocal = rowScale[j:j + nG, c + 1].clone(), local[sl, c + 1].clone()
                eng.run('backward', plan, _engine.F_RAW_POSTERIOR, T=c + 2, B=nG, program=sw['program'], lo=s0 + j,
                        alpha_src=alphaS, src_stride=T * G,  # out of place: the group's filtering rows are read in place
                        alpha_seq=buf[j:j + nG], seq_stride=T * G, row_scale=rowScale[j:j + nG], row_stride=T,
                        local_evidence=local[sl], log_evidence=logE[sl], alive=alive[sl], **base, **window(0))
                buf[j:j + nG, c + 1] = sw['saveRow']
                rowScale[j:j + nG, c + 1] = keepScale
                local[sl, c + 1] = keepLocal
                executed += nG * (c + 2)
            eng.run('accumulate', plan, 0, T=T, B=nb, program=sw['program'], lo=s0, log_weight=weights[s0:s0 + nb], avg=avg,
                    alpha_seq=buf, row_scale=rowScale, alive=alive[s0:s0 + nb], log_evidence=logE[s0:s0 + nb], **base,
                    **window(0))
        if not bool((alive[:B] == 1).all()):
            # a combination died on the way (zero norm, core.py:388-400, :440-452): the reference leaves the local
            # evidences of its forward pass in the rows the backward pass never reached (core.py:1356 keeps them) --
            # the plain sweep reproduces exactly that, the windowed passes do not
            return None
        part = sw['part']
        part.zero_()
        eng.mix(plan, local, sw['hpDev'], B, T, part)
        localEv = dist.reduce_sum(eng, part)
        dist.rebase_and_reduce(eng, plan, avg, shift, T * G)
        eng.finalize(plan, avg, T, sw['means'], _engine.F_NORMALIZE_ROWS)
        logEAll, aliveAll = dist.gather_rows(eng, logE[:B], alive[:B], sw['Ball'], rows=sw['rows'])
        logEAll = np.where(aliveAll == 1, logEAll, -np.inf)
        self.sweepStats = dict(waves=waves, wave=sw['wave'], rows=sw['rows'], launches=eng.launch_count(), shared=True,
                               deal=share.get('deal'), executed_updates=int(executed) * G, nominal_updates=2 * int(B) * T * G)
        return eng, logEAll, aliveAll, localEv, avg, sw['means']

    def _executeSweep(self, sw):
        """Kernels of one sweep (inputs already resident).  Per wave: forward pass; the wave's averaging weights and
        the re-base of the running sum are computed ON THE DEVICE from the evidences (blg_wave_weights: the streaming
        form of np.logaddexp, core.py:1358-1366); backward pass; weighted accumulation.  Everything is enqueued on
        one stream without a host round trip: the host reads the evidences once, after the cross-rank merge."""
        from . import distributed as dist
        if sw.get('share') is not None:
            out = self._executeSharedSweep(sw)
            if out is not None:
                return out
        eng, ses, T, G, B = sw['eng'], sw['ses'], sw['T'], sw['G'], sw['B']
        evidenceOnly, forwardOnly = sw['evidenceOnly'], sw['forwardOnly']
        logE, local, alive, avg, buf, wave = sw['logE'], sw['local'], sw['alive'], sw['avg'], sw['buf'], sw['wave']
        shift, weights = sw['shift'], sw['weights']
        if avg is not None:
            avg.zero_()
        shift.fill_(-math.inf)
        waves = 0
        for w0 in range(0, B, wave):
            w1 = min(B, w0 + wave)
            nb = w1 - w0
            common = dict(T=T, B=nb, data=ses.data, prior=ses.prior, lik_table=ses.likTable, program=sw['program'],
                          lo=w0, reset_base=sw['resetBase'], log_evidence=logE[w0:w1], local_evidence=local[w0:w1],
                          alive=alive[w0:w1], alpha_seq=buf)
            eng.run('forward', ses.plan,
                    _engine.F_EVIDENCE_ONLY if evidenceOnly else (0 if forwardOnly else _engine.F_RAW_ALPHA), **common)
            waves += 1
            if evidenceOnly:
                continue
            # the evidences of the wave fix its averaging weights; a larger reference log-weight re-bases the sum
            eng.wave_weights(ses.plan, logE[w0:w1], sw['logHpDev'][w0:w1], nb, shift, avg, T * G, weights[w0:w1])
            if not forwardOnly:
                # smoothed posteriors overwrite the stored sequences in place; combos whose backward pass hits a
                # zero norm flag themselves (alive = -1) and are dropped from the average as a whole (core.py:1358)
                # (rows may come back unnormalised with their scale in rowScale, applied by the accumulation)
                eng.run('backward', ses.plan, _engine.F_RAW_POSTERIOR, row_scale=sw['rowScale'], **common)
            # evidence-weighted sum over the combos of this wave, fixed summation order (deterministic)
            eng.run('accumulate', ses.plan, 0, log_weight=weights[w0:w1], avg=avg,
                    row_scale=None if forwardOnly else sw['rowScale'], **common)
        # averaged local evidence: sum_b localEvidence_b * hyperprior_b (core.py:1410)
        part = sw['part']
        part.zero_()
        if B > 0:
            eng.mix(ses.plan, local, sw['hpDev'], B, T, part)
        # cross-rank merge, all collectives enqueued back to back; ONE device -> host read of the evidences at the end
        localEv = dist.reduce_sum(eng, part)
        if not evidenceOnly:
            dist.rebase_and_reduce(eng, ses.plan, avg, shift, T * G)
            eng.finalize(ses.plan, avg, T, sw['means'], _engine.F_NORMALIZE_ROWS)
        logEAll, aliveAll = dist.gather_rows(eng, logE[:B], alive[:B], sw['Ball'], rows=sw['rows'])
        logEAll = np.where(aliveAll == 1, logEAll, -np.inf)
        self.sweepStats = dict(waves=waves, wave=wave, rows=sw['rows'], launches=eng.launch_count(), shared=False,
                               executed_updates=(1 if (evidenceOnly or forwardOnly) else 2) * int(B) * T * G,
                               nominal_updates=(1 if (evidenceOnly or forwardOnly) else 2) * int(B) * T * G)
        return eng, logEAll, aliveAll, localEv, avg, sw['means']

    def _sweep(self, forwardOnly, evidenceOnly):
        sw = self._prepareSweep(forwardOnly, evidenceOnly)
        self._sweepPlan = sw['ses'].plan
        eng, logE, alive, localEv, avg, means = self._executeSweep(sw)
        return eng, logE, alive, eng.to_host(localEv), avg, means

    @property
    def averagePosteriorSequence(self):
        """Evidence-weighted average of the posterior sequences (core.py:1372-1385): the fitted sequence itself."""
        if self.__dict__.get('_avgIsPosterior'):
            return self.posteriorSequence
        return self.__dict__.get('_avgHost')

    @averagePosteriorSequence.setter
    def averagePosteriorSequence(self, value):
        self.__dict__['_avgIsPosterior'] = False
        self.__dict__['_avgHost'] = value

    def fit(self, forwardOnly=False, evidenceOnly=False, silent=False, nJobs=1, customHyperGrid=False):
        """Fit every combination of hyper-parameter values and average the models by their evidence (contract of
        core.py:1247-1264).  `nJobs` is accepted for compatibility; device parallelism comes from the batch
        kernels and from torch.distributed ranks (one per GPU), not from a process pool."""
        self.fitWarningCounter = 0
        self._formatData()
        if not customHyperGrid:
            self._createHyperGrid(silent=silent)
            points = list(flatten(self._unpackChangepointNames(self.transitionModel))) + \
                list(flatten(self._unpackBreakpointNames(self.transitionModel)))
            if len(points) > 1:
                cols = [self.flatHyperParameterNames.index(p) for p in points]
                sub = np.sort(np.asarray(self.hyperGridValues)[:, cols], axis=1)
                if np.any(np.diff(sub, axis=1) == 0):
                    raise ConfigurationError('Detected multiple change-/break-points with identical values and/or '
                                             'overlapping value intervals. Use "ChangepointStudy" instead of '
                                             '"HyperStudy" for such cases.')
        self._checkConsistency()
        self.logEvidenceList = []
        self.localEvidenceList = []

        if len(self.hyperGridValues) <= 1:
            if not silent:
                if len(self.hyperGridValues) == 1:
                    print('+ Only one combination of hyper-parameter values, switching to standard fit method.')
                else:
                    print('+ Transition model contains no hyper-parameters, switching to standard fit method.')
            if not evidenceOnly:
                self.averagePosteriorSequence = None
            if len(self.hyperGridValues) == 1:  # single-element lists -> scalars for the duration of the fit
                self._setAllHyperParameters(list(self.hyperGridValues[0]))
                try:
                    Study.fit(self, forwardOnly=forwardOnly, evidenceOnly=evidenceOnly, silent=silent)
                finally:
                    self._setAllHyperParameters(self.flatHyperParameters)
            else:
                Study.fit(self, forwardOnly=forwardOnly, evidenceOnly=evidenceOnly, silent=silent)
            return

        if len(self.formattedData) == 0:
            # the reference fails in its averaging step with this exception type (core.py:1372, np.amax of an empty array)
            raise ValueError('zero-size series: fewer data points than one segment of the observation model')
        if not silent:
            print('+ Started new fit.')
            print('    + {} analyses to run.'.format(len(self.hyperGridValues)))
        eng, logE, alive, localEv, avg, means = self._sweep(forwardOnly, evidenceOnly)
        for state in alive[alive != 1][:6]:
            self._warnZeroNorm(backward=(state == -1))
        self.logEvidenceList = list(logE)
        T = len(self.formattedData)
        if not evidenceOnly:
            # the averaged sequence stays in HBM until somebody reads it (posteriorSequence / averagePosteriorSequence)
            self._setDeviceSequence(eng, self._sweepPlan, avg, [T] + list(self.gridSize))
            self.__dict__['_avgIsPosterior'] = True
            if not silent:
                print('    + Computed average posterior sequence')

        # hyper-parameter distribution and evidence of the average model (core.py:1391-1410)
        with np.errstate(divide='ignore'):
            logDist = logE + np.log(self.flatHyperPriorValues) + np.sum(np.log(self.hyperGridConstant))
        top = np.amax(logDist)
        dist = np.exp(logDist - top)
        self.hyperParameterDistribution = dist / np.sum(dist) / np.prod(self.hyperGridConstant)
        # scipy's logsumexp (core.py:1405) hands back a non-finite maximum as it is: -inf (every combination dead), NaN, +inf
        self.logEvidence = float(top + np.log(np.sum(dist))) if np.isfinite(top) else float(top)
        self.localEvidence = localEv
        if not silent:
            print('    + Computed hyper-parameter distribution')
            print('    + Log10-evidence of average model: {:.5f}'.format(self.logEvidence / np.log(10)))
            print('    + Computed local evidence of average model')
        if not evidenceOnly:
            self.posteriorMeanValues = eng.to_host(means)
            if not silent:
                print('    + Computed mean parameter values.')
        self.localEvidenceList = []
        self._setAllHyperParameters(self.flatHyperParameters)
        if not silent:
            print('+ Finished fit.')

    def optimize(self, *args, **kwargs):
        raise NotImplementedError('HyperStudy object has no optimizing method.')

    # ------------------------------------------------------------------------------------------- accessors
    def _hyperShape(self):
        return [len(x) if isinstance(x, Iterable) else 1 for x in self.flatHyperParameters]

    def getHyperParameterDistribution(self, name, plot=False, **kwargs):
        """Marginal probability of one hyper-parameter over its value grid."""
        if len(self.hyperGridValues) < 2:
            raise PostProcessingError('At least two combinations of hyper-parameter values need to be fitted to '
                                      'evaluate a hyper-parameter distribution. Check transition model.')
        axis = self._getHyperParameterIndex(self.transitionModel, name)
        cube = np.asarray(self.hyperParameterDistribution).reshape(self._hyperShape(), order='C')
        others = tuple(a for a in range(cube.ndim) if a != axis)
        marginal = (np.sum(cube, axis=others) if others else cube) * np.prod(self.hyperGridConstant)
        return self.flatHyperParameters[axis], marginal

    def getHPD(self, name, plot=False, **kwargs):
        return self.getHyperParameterDistribution(name, plot=plot, **kwargs)

    def getJointHyperParameterDistribution(self, names, plot=False, figure=None, subplot=111, **kwargs):
        """Joint probability of two hyper-parameters (rows follow names[0], columns names[1])."""
        if len(self.hyperGridValues) < 2:
            raise PostProcessingError('At least two combinations of hyper-parameter values need to be fitted to '
                                      'evaluate a hyper-parameter distribution. Check transition model.')
        if not isinstance(names, Iterable) or len(names) != 2:
            raise PostProcessingError('A list of exactly two hyper-parameters has to be provided.')
        a, b = [self._getHyperParameterIndex(self.transitionModel, n) for n in names]
        cube = np.asarray(self.hyperParameterDistribution).reshape(self._hyperShape(), order='C')
        others = tuple(ax for ax in range(cube.ndim) if ax not in (a, b))
        joint = (np.sum(cube, axis=others) if others else cube) * np.prod(self.hyperGridConstant)
        if a > b:
            joint = joint.T
        return self.flatHyperParameters[a], self.flatHyperParameters[b], joint

    def getJHPD(self, names, plot=False, figure=None, subplot=111, **kwargs):
        return self.getJointHyperParameterDistribution(names, plot=plot, figure=figure, subplot=subplot, **kwargs)


class ChangepointStudy(HyperStudy):
    """Sweep over all ORDERED combinations of change-/break-point times (reference: core.py:1743-1930)."""

    def __init__(self, silent=False, engine=None):
        super(ChangepointStudy, self).__init__(silent=silent, engine=engine)
        self.allHyperGridValues = []
        self.allHyperPriorValues = []
        self.mask = []
        self.userDefinedGrid = False
        self.hyperGridBackup = []
        if not silent:
            print('  --> Change-point analysis')

    def _unpackSerialTransitionModels(self, transitionModel):
        out = [self._unpackSerialTransitionModels(m) for m in getattr(transitionModel, 'models', [])]
        if hasattr(transitionModel, 'hyperParameterNames') and str(transitionModel) == 'Serial transition model':
            out.append(transitionModel)
        return out

    def _prepareChangepoints(self, silent=False):
        """Hyper-grid of a change-point study: Cartesian grid, then only the strictly ordered tuples of change-/
        break-point times are kept and the prior mass of the full grid is restored (core.py:1777-1834)."""
        self._formatData()
        if len(list(flatten(self._unpackSerialTransitionModels(self.transitionModel)))) > 1:
            raise NotImplementedError('Multiple instances of SerialTransition models are currently not supported by '
                                      'ChangepointStudy.')
        changepoints = list(flatten(self._unpackChangepointNames(self.transitionModel)))
        breakpoints = list(flatten(self._unpackBreakpointNames(self.transitionModel)))
        if changepoints and breakpoints:
            raise NotImplementedError('Detected both change-points (Changepoint transition model) and break-points '
                                      '(SerialTransitionModel). Currently, only one type is supported in a single '
                                      'transition model.')
        if not changepoints and not breakpoints:
            raise ConfigurationError('No change-points or break-points detected in transition model. Check transition '
                                     'model.')
        self.flatHyperParameters = self._unpackAllHyperParameters()
        self.flatHyperParameterNames = self._unpackAllHyperParameters(values=False)
        points = changepoints if changepoints else breakpoints
        if not silent:
            print('+ Detected {} {}-point(s) in transition model: {}'
                  .format(len(points), 'change' if changepoints else 'break', points))

        self._createHyperGrid(silent=silent)
        self.allHyperGridValues = self.hyperGridValues[:]
        self.allHyperPriorValues = self.flatHyperPriorValues[:]
        cols = np.isin(np.array(self.flatHyperParameterNames), points)
        times = self.allHyperGridValues[:, cols]
        # keep strictly increasing tuples only (core.py:1826-1834), then restore the prior mass of the full grid
        self.mask = np.all(times[:, :-1] < times[:, 1:], axis=1) if times.shape[1] > 1 else \
            np.ones(len(times), dtype=bool)
        self.hyperGridValues = self.allHyperGridValues[self.mask]
        self.flatHyperPriorValues = self.allHyperPriorValues[self.mask]
        self.flatHyperPriorValues = self.flatHyperPriorValues * (np.sum(self.allHyperPriorValues) /
                                                                 np.sum(self.allHyperPriorValues[self.mask]))

    def fit(self, forwardOnly=False, evidenceOnly=False, silent=False, nJobs=1):
        self._prepareChangepoints(silent=silent)
        HyperStudy.fit(self, forwardOnly=forwardOnly, evidenceOnly=evidenceOnly, silent=silent, nJobs=nJobs,
                       customHyperGrid=True)
        full = np.zeros(len(self.allHyperGridValues))
        full[self.mask] = self.hyperParameterDistribution
        self.hyperParameterDistribution = full
        full = np.zeros(len(self.allHyperPriorValues))
        full[self.mask] = self.flatHyperPriorValues
        self.flatHyperPriorValues = full

    def getDurationDistribution(self, names, plot=False, **kwargs):
        """Distribution of the time between two change-/break-points (reference: core.py:1875-1924)."""
        if not isinstance(names, Iterable) or len(names) != 2:
            raise PostProcessingError('A list of exactly two hyper-parameters has to be provided.')
        a, b = sorted(self._getHyperParameterIndex(self.transitionModel, n) for n in names)
        allValues = np.asarray(self.allHyperGridValues)
        delta = allValues[:, b] - allValues[:, a]
        keep = delta > 0
        duration = np.unique(np.asarray(self.hyperGridValues)[:, b] - np.asarray(self.hyperGridValues)[:, a])
        idx = np.searchsorted(duration.round(10), delta[keep].round(10))
        out = np.zeros(len(duration))
        np.add.at(out, idx, np.asarray(self.hyperParameterDistribution)[keep])
        return duration, out / np.sum(out)

    def getDD(self, names, plot=False, **kwargs):
        return self.getDurationDistribution(names, plot=plot, **kwargs)


class OnlineStudy(HyperStudy):
    """Streaming forward filter over many transition-model hypotheses (reference: core.py:1933-2226).

    All hypotheses -- every hyper-parameter combination of every added transition model -- live in ONE device
    batch: their programs are concatenated and masked per hypothesis, the per-hypothesis posteriors stay resident
    in HBM between `step` calls, and each `step` is one batched kernel launch plus O(#hypotheses) host arithmetic.

    Under torch.distributed (one process per GPU) the hypotheses are dealt round-robin over the ranks
    (distributed.py): every rank calls `step` with the same data point and ends up with identical evidences and
    distributions; `marginalizedPosterior`, `parameterPosterior` and `transitionModelPosterior` are then COLLECTIVE
    reads (all-reduce / all-gather), to be made by all ranks together.
    """

    def __init__(self, storeHistory=False, silent=False, engine=None):
        super(OnlineStudy, self).__init__(silent=silent, engine=engine)
        self.firstStep = True
        self.transitionModels = []
        self.transitionModelNames = []
        self.tmCount = None
        self.tmCounts = []
        self.hyperParameterValues = []
        self.allFlatHyperParameterValues = []
        self.hyperParameterNames = []
        self.hyperGridConstants = []
        self.logEvidenceList = None
        self.hyperLogEvidenceList = None
        self.hyperPrior = []
        self.hyperPriorValues = []
        self.transitionModelPrior = None
        self.hyperParameterDistribution = None
        self.transitionModelDistribution = None
        self.localTransitionModelDistribution = None
        self.storeHistory = storeHistory
        self.posteriorMeanValues = []
        self.posteriorSequence = []
        self.hyperParameterSequence = []
        self.transitionModelSequence = []
        self.localTransitionModelSequence = []
        self._dev = None
        if not silent:
            print('  --> Online study')

    def addTransitionModel(self, name, transitionModel):
        self.setTransitionModel(transitionModel, silent=True)
        self._createHyperGrid(silent=True)
        self.transitionModels.append(transitionModel)
        self.transitionModelNames.append(name)
        self.hyperParameterValues.append(self.hyperGridValues[:])
        self.allFlatHyperParameterValues.append(self.flatHyperParameters)
        self.hyperParameterNames.append(self.flatHyperParameterNames[:])
        self.hyperGridConstants.append(self.hyperGridConstant[:])
        self.hyperPrior.append(self.flatHyperPriors[:])
        self.hyperPriorValues.append(self.flatHyperPriorValues[:])
        self.tmCounts = [len(h) if len(h) > 0 else 1 for h in self.hyperParameterValues]
        self.tmCount = int(np.sum(self.tmCounts))
        if len(self.hyperGridValues) > 0:
            print('+ Added transition model: {} ({} combination(s) of the following hyper-parameters: {})'
                  .format(name, len(self.hyperGridValues), self.hyperParameterNames[-1]))
        else:
            print('+ Added transition model: {} (no hyper-parameters)'.format(name))

    def addTM(self, name, transitionModel):
        self.addTransitionModel(name, transitionModel)

    def add(self, name, transitionModel):
        self.addTransitionModel(name, transitionModel)

    def setTransitionModelPrior(self, transitionModelPrior, silent=False):
        if not (isinstance(transitionModelPrior, Iterable) and
                len(transitionModelPrior) == len(self.transitionModels)):
            raise ConfigurationError('Length of transition model prior ({}) does not fit number of transition models '
                                     '({})'.format(len(transitionModelPrior), len(self.transitionModels)))
        self.transitionModelPrior = np.array(transitionModelPrior, dtype=float)
        if not np.sum(transitionModelPrior) == 1.:
            print('+ WARNING: Transition model prior does not sum up to one. Will re-normalize.')
            self.transitionModelPrior /= np.sum(self.transitionModelPrior)
        if not silent:
            print('+ Set custom transition model prior.')

    def _setupDevice(self):
        """Concatenate the programs of all transition models; hypothesis h only sees the operators of its own
        model (all other windows are empty).  t = -1 is what the reference hands to the models (core.py:2167)."""
        eng = self._engine()
        om = self.observationModel
        seg = om.segmentLength
        first = np.asarray(self.rawData[-seg:], dtype=float)
        nCols = 1 if first.ndim == 1 else int(first.shape[1])
        kind = getattr(om, 'deviceKind', KIND_TABLE)
        if type(om).pdf is not ObservationModel.pdf:
            kind = KIND_TABLE
        plan = eng.plan(self.marginalGrid, self.latticeConstant, kind, seg, nCols)
        H = self.tmCount
        ops, usesReset, row = [], False, 0
        for tm, rows, count in zip(self.transitionModels, self.hyperParameterValues, self.tmCounts):
            hyper = np.asarray(rows, dtype=float).reshape(count, -1) if len(rows) > 0 else np.zeros((1, 0))
            self.setTransitionModel(tm, silent=True)
            ctx = self._lower(hyper, np.array([-1.]), online=True, model=tm)
            usesReset |= ctx.usesReset
            for op in ctx.ops:
                full = dict(kind=op['kind'], axis=op['axis'], param=np.zeros(H), radius=np.zeros(H, dtype=np.int32),
                            window=np.zeros((H, 4), dtype=np.int32))
                full['param'][row:row + count] = op['param']
                full['radius'][row:row + count] = op['radius']
                full['window'][row:row + count] = op['window']
                ops.append(full)
            row += count
        G = int(np.prod(self.gridSize))
        # one process per GPU: hypothesis h lives on rank h % world (distributed.py); a single process owns them all
        from . import distributed as dist
        rows = dist.shard_rows(H)
        Hr = len(rows)
        for op in ops:
            op['param'], op['radius'], op['window'] = op['param'][rows], op['radius'][rows], op['window'][rows]
        dev = dict(eng=eng, plan=plan, kind=kind, nCols=nCols, G=G, H=H, rows=rows, Hr=Hr,
                   separable=_rows_separable(ops, Hr),
                   program=_engine.Program(eng, ops, Hr), state=eng.empty((Hr, G)), step=eng.zeros(Hr),
                   alive=eng.zeros(Hr, dtype=torch.int32), mixed=eng.empty(G), tmPost=None,
                   prior=eng.to_device(np.asarray(self._computePrior(silent=False), dtype=float).reshape(-1)),
                   resetBase=None)
        if usesReset:
            dev['resetBase'] = _Session.reset_base(_ResetShim(self, eng))
        self._dev = dev

    def step(self, dataPoint):
        """Include one new data point (contract of core.py:2062-2226)."""
        if self.tmCount is None and self.transitionModel is None:
            raise ConfigurationError('No transition model set or added.')
        if self.tmCount is None:
            self.addTransitionModel('transition model', self.transitionModel)
        if not isinstance(dataPoint, list):
            dataPoint = [dataPoint]
        if len(self.rawData) == 0:
            print('+ Start model fit')
            names = list(flatten(self.hyperParameterNames))
            if len(names) != len(set(names)):
                raise ConfigurationError('Detected duplicate hyper-parameter names. Choose unique identifiers.')
            self.rawData = np.array(dataPoint)
            Study._checkConsistency(self)
            self.rawTimestamps = np.array([0])
            self.formattedTimestamps = []
        else:
            self.rawData = np.append(self.rawData, np.array(dataPoint), axis=0)
            self.rawTimestamps = np.append(self.rawTimestamps, self.rawTimestamps[-1] + 1)
        seg = self.observationModel.segmentLength
        if len(self.rawData) < seg:
            print('+ Not enough data points to start analysis. Will wait for more data.')
            return
        self.formattedTimestamps.append(self.rawTimestamps[-1])

        nTM = len(self.transitionModels)
        self._resume()
        if self.firstStep:
            self._setupDevice()
            if self.transitionModelPrior is None:
                self.transitionModelPrior = np.ones(nTM) / nTM
                if nTM > 1:
                    print('    + Set flat transition model prior.')
            self.logEvidenceList = [np.zeros(c) for c in self.tmCounts]
            self.hyperLogEvidenceList = np.zeros(nTM)
            self.hyperParameterDistribution = [np.zeros(c) for c in self.tmCounts]
            self.transitionModelDistribution = np.zeros(nTM)
            self.localTransitionModelDistribution = np.zeros(nTM)
        dev = self._dev
        eng, plan, H, G = dev['eng'], dev['plan'], dev['H'], dev['G']

        segment = np.asarray(self.rawData[-seg:], dtype=float).reshape(seg, dev['nCols'])
        likTable = None
        if dev['kind'] == KIND_TABLE:
            seg1 = self.rawData[-seg:]
            likTable = eng.to_device(np.asarray(self.observationModel.processedPdf(self.grid, seg1),
                                                dtype=float).reshape(1, G))
        flags = _engine.F_EVIDENCE_ONLY | _engine.F_SAVE_STATE
        if not self.firstStep:
            flags |= _engine.F_INIT_STATE | _engine.F_TRANSITION_FIRST
        if dev['separable']:
            flags |= _engine.F_SEPARABLE_ROWS
        if dev['Hr'] > 0:
            eng.run('forward', plan, flags, T=1, B=dev['Hr'], data=eng.to_device(segment), prior=dev['prior'],
                    reset_base=dev['resetBase'], lik_table=likTable, program=dev['program'],
                    init_state=dev['state'], log_evidence=dev['step'], alive=dev['alive'], final_state=dev['state'])
        # log n_i per hypothesis (+ log prod(latticeConstant) on the first step); with several ranks this all-gather
        # of H doubles is the only per-step exchange
        from . import distributed as dist
        inc = dist.gather_dealt(eng, dev['step'], H)

        # O(H) bookkeeping of core.py:2171-2215
        weights = np.zeros(H)
        row = 0
        for i, count in enumerate(self.tmCounts):
            self.logEvidenceList[i] = self.logEvidenceList[i] + inc[row:row + count]
            with np.errstate(divide='ignore'):
                logPost = self.logEvidenceList[i] + np.log(self.hyperPriorValues[i])
            old = self.hyperLogEvidenceList[i]
            top = np.amax(logPost)
            self.hyperLogEvidenceList[i] = top + np.log(np.sum(np.exp(logPost - top)))
            self.transitionModelDistribution[i] = self.hyperLogEvidenceList[i]
            self.localTransitionModelDistribution[i] = self.hyperLogEvidenceList[i] - old + \
                np.log(self.transitionModelPrior[i])
            hpd = np.exp(logPost - top)
            hpd /= np.sum(hpd)
            if len(self.hyperGridConstants[i]) > 0:
                hpd /= np.prod(self.hyperGridConstants[i])
            self.hyperParameterDistribution[i] = hpd
            weights[row:row + count] = hpd * np.prod(self.hyperGridConstants[i])
            row += count
        for attr in ('transitionModelDistribution', 'localTransitionModelDistribution'):
            d = getattr(self, attr)
            d = np.exp(d - np.amax(d))
            setattr(self, attr, d / np.sum(d))
        self.logEvidence = float(_logsumexp(self.hyperLogEvidenceList + np.log(self.transitionModelPrior)))

        # marginalisation over hyper-parameters and transition models: one weighted row-sum on the device
        row = 0
        for i, count in enumerate(self.tmCounts):
            weights[row:row + count] *= self.transitionModelDistribution[i]
            row += count
        dev['weights'] = weights
        dev['mixedValid'] = False
        if self.storeHistory:
            post = self.marginalizedPosterior
            self.posteriorMeanValues.append(np.array([np.sum(post * g) for g in self.grid]))
            self.posteriorSequence.append(post.copy())
            self.hyperParameterSequence.append([h.copy() for h in self.hyperParameterDistribution])
            self.transitionModelSequence.append(self.transitionModelDistribution.copy())
            self.localTransitionModelSequence.append(self.localTransitionModelDistribution.copy())
        self.firstStep = False

    # ------------------------------------------------------------------------------------ checkpoint / resume
    # SURVEY.md 8f row f3 (reference: bl.save / bl.load pickle the whole study, fileIO.py:10-37).  The hypotheses'
    # posteriors live in HBM between steps; pickling brings them to the host, unpickling re-creates the device side
    # lazily (plan, concatenated program, state upload) on whatever engine is current -- a study checkpointed on one
    # GPU resumes on another, or on a different rank.
    def __getstate__(self):
        d = self.__dict__.copy()
        dev = d.pop('_dev', None)
        d['_dev'] = None
        d['_engineOverride'] = None  # engines hold library handles and a device: not part of the state
        if dev is not None:
            eng = dev['eng']
            d['_checkpoint'] = dict(state=eng.to_host(dev['state']).copy(), weights=np.array(dev.get('weights')),
                                    H=dev['H'], G=dev['G'], rows=np.array(dev['rows']))
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._attachBackPointers()

    def _resume(self):
        """Re-create the device side of an unpickled study (no-op otherwise)."""
        ck = self.__dict__.get('_checkpoint')
        if self._dev is None and ck is not None and not self.firstStep:
            self._setupDevice()
            dev = self._dev
            if (dev['H'], dev['G']) != (ck['H'], ck['G']):
                raise ConfigurationError('Checkpoint does not match the models of this study.')
            if not np.array_equal(ck.get('rows', np.arange(ck['H'])), dev['rows']):
                raise ConfigurationError('Checkpoint was written by a different rank / world size: each rank resumes '
                                         'the hypotheses it owned.')
            dev['state'].copy_(dev['eng'].to_device(ck['state']))
            dev['weights'] = ck['weights']
            dev['mixedValid'] = False
            self._checkpoint = None

    # device-resident results, copied to the host when somebody looks at them
    @property
    def marginalizedPosterior(self):
        self._resume()
        dev = self._dev
        if dev is None:
            return None
        if not dev.get('mixedValid'):
            eng = dev['eng']
            from . import distributed as dist
            if dev['Hr'] > 0:
                eng.mix(dev['plan'], dev['state'], eng.to_device(dev['weights'][dev['rows']]), dev['Hr'], dev['G'],
                        dev['mixed'])
            else:
                dev['mixed'].zero_()
            dist.reduce_sum(eng, dev['mixed'])  # per-rank weighted row sums -> the mixture over all hypotheses
            dev['mixedHost'] = eng.to_host(dev['mixed']).reshape(self.gridSize)
            dev['mixedValid'] = True
        return dev['mixedHost']

    @marginalizedPosterior.setter
    def marginalizedPosterior(self, value):
        pass

    @property
    def parameterPosterior(self):
        self._resume()
        dev = self._dev
        if dev is None:
            return None
        from . import distributed as dist
        flat = dist.gather_dealt(dev['eng'], dev['state'], dev['H'])
        out, row = [], 0
        for count in self.tmCounts:
            out.append(flat[row:row + count].reshape([count] + self.gridSize))
            row += count
        return out

    @parameterPosterior.setter
    def parameterPosterior(self, value):
        pass

    @property
    def transitionModelPosterior(self):
        posts = self.parameterPosterior
        if posts is None:
            return None
        out = np.zeros([len(posts)] + self.gridSize)
        for i, (p, hpd) in enumerate(zip(posts, self.hyperParameterDistribution)):
            w = (hpd * np.prod(self.hyperGridConstants[i])).reshape([-1] + [1] * len(self.gridSize))
            out[i] = np.sum(p * w, axis=0)
        return out

    @transitionModelPosterior.setter
    def transitionModelPosterior(self, value):
        pass

    def fit(self, *args, **kwargs):
        raise NotImplementedError('OnlineStudy object has no "fit" method. Use "step" instead.')

    # ------------------------------------------------------------------------------------------- accessors
    # Host-side views of the streaming results (reference: core.py:2231-2897).  The "current" ones read the device
    # state of the last step; the time-indexed ones need storeHistory=True (the history lives on the host).
    def _needHistory(self, what, instead):
        if not self.storeHistory:
            raise PostProcessingError('To get past {}, Online Study must be called with flag "storeHistory=True". '
                                      'Use "{}" instead.'.format(what, instead))

    def _timeIndex(self, t):
        stamps = list(self.formattedTimestamps)
        if t not in stamps:
            raise PostProcessingError('Supplied time ({}) does not exist in data or is out of range.'.format(t))
        return stamps.index(t)

    def _locateHyperParameter(self, name):
        """(transition-model index, hyper-parameter index) of a hyper-parameter of any added model; the last
        model that knows the name wins, like the loop of core.py:2593-2599."""
        found = None
        for i, tm in enumerate(self.transitionModels):
            try:
                found = (i, self._getHyperParameterIndex(tm, name))
            except PostProcessingError:
                pass
        if found is None:
            raise PostProcessingError('No hyper-parameter "{}" found. Check hyper-parameter names.'.format(name))
        return found

    def _hyperMarginal(self, flat, tmIndex, hpIndex):
        steps = [len(x) for x in self.allFlatHyperParameterValues[tmIndex]]
        cube = np.asarray(flat, dtype=float).reshape(steps, order='C')
        others = tuple(a for a in range(cube.ndim) if a != hpIndex)
        return np.sum(cube, axis=others) if others else cube.copy()

    def getParameterDistribution(self, t, name, plot=False, density=True, **kwargs):
        """Marginal distribution of one parameter at time stamp `t` or 'avg' (core.py:2231-2260)."""
        self._needHistory('parameter distributions', 'getCurrentParameterDistribution')
        seq = np.asarray(self.posteriorSequence)
        dist = np.sum(seq, axis=0) / len(seq) if isinstance(t, str) and t == 'avg' else seq[self._timeIndex(t)]
        axis = self._parameterIndex(name)
        others = tuple(a for a in range(dist.ndim) if a != axis)
        marginal = np.sum(dist, axis=others) if others else dist.copy()
        if density:
            marginal = marginal / self.latticeConstant[axis]
        return self.marginalGrid[axis], marginal

    def getPD(self, t, name, plot=False, density=True, **kwargs):
        return self.getParameterDistribution(t, name, plot=plot, density=density, **kwargs)

    def getCurrentParameterDistribution(self, name, plot=False, density=True, **kwargs):
        """Marginal distribution of one parameter after the last step (core.py:2268-2315)."""
        axis = self._parameterIndex(name)
        post = self.marginalizedPosterior
        others = tuple(a for a in range(post.ndim) if a != axis)
        marginal = np.sum(post, axis=others) if others else post.copy()
        if density:
            marginal = marginal / self.latticeConstant[axis]
        return self.marginalGrid[axis], marginal

    def getCPD(self, name, plot=False, density=True, **kwargs):
        return self.getCurrentParameterDistribution(name, plot=plot, density=density, **kwargs)

    def getParameterDistributions(self, name, plot=False, density=True, **kwargs):
        """Marginal distributions of one parameter for all steps so far: [T, n_axis] (core.py:2323-2351)."""
        self._needHistory('parameter distributions', 'getCurrentParameterDistribution')
        axis = self._parameterIndex(name)
        if self.__dict__.get('_postDev') is not None:  # sequence still in HBM: only [T x n_axis] numbers come back
            marginal = self._deviceMarginal(axis)
        else:
            seq = np.asarray(self.posteriorSequence)
            others = tuple(a + 1 for a in range(seq.ndim - 1) if a != axis)
            marginal = np.sum(seq, axis=others) if others else seq.copy()
        if density:
            marginal = marginal / self.latticeConstant[axis]
        return self.marginalGrid[axis], marginal

    def getPDs(self, name, plot=False, density=True, **kwargs):
        return self.getParameterDistributions(name, plot=plot, density=density, **kwargs)

    def getCurrentTransitionModelDistribution(self, local=False, plot=False, **kwargs):
        dist = self.localTransitionModelDistribution if local else self.transitionModelDistribution
        return np.array(self.transitionModelNames), dist

    def getCTMD(self, local=False):
        return self.getCurrentTransitionModelDistribution(local=local)

    def getCurrentTransitionModelProbability(self, transitionModel, local=False):
        i = self.transitionModelNames.index(transitionModel)
        return (self.localTransitionModelDistribution if local else self.transitionModelDistribution)[i]

    def getCTMP(self, transitionModel, local=False):
        return self.getCurrentTransitionModelProbability(transitionModel, local=local)

    def getTransitionModelDistributions(self, local=False):
        """Names and [T, #models] probabilities of the transition models for all steps (core.py:2430-2450)."""
        self._needHistory('transition model distributions', 'getCurrentTransitionModelDistribution')
        seq = self.localTransitionModelSequence if local else self.transitionModelSequence
        return np.array(self.transitionModelNames), np.array(seq)

    def getTransitionModelProbabilities(self, transitionModel, local=False):
        self._needHistory('transition model distributions', 'getCurrentTransitionModelDistribution')
        i = self.transitionModelNames.index(transitionModel)
        seq = self.localTransitionModelSequence if local else self.transitionModelSequence
        return np.array(seq)[:, i]

    def getTMPs(self, transitionModel, local=False):
        return self.getTransitionModelProbabilities(transitionModel, local=local)

    def getCurrentParameterMeanValue(self, name):
        return np.sum(self.marginalizedPosterior * self.grid[self._parameterIndex(name)])

    def getParameterMeanValue(self, t, name):
        """Posterior mean of one parameter at time stamp `t`.  (The reference indexes the stored posterior a second
        time with `t`, core.py:2541 -- a quirk that only works on 1-D grids by accident; this is the mean of the
        whole posterior of that step, i.e. getParameterMeanValues(name)[index of t].)"""
        self._needHistory('parameter mean values', 'getCurrentParameterMeanValue')
        axis = self._parameterIndex(name)
        return np.sum(np.asarray(self.posteriorSequence[self._timeIndex(t)]) * self.grid[axis])

    def getParameterMeanValues(self, name):
        self._needHistory('parameter mean values', 'getCurrentParameterMeanValue')
        return np.array(self.posteriorMeanValues).T[self._parameterIndex(name)]

    def _hyperMean(self, flat, tmIndex, hpIndex):
        values = np.asarray(self.hyperParameterValues[tmIndex], dtype=float)[:, hpIndex]
        return np.sum(values * np.asarray(flat, dtype=float)) * np.prod(self.hyperGridConstants[tmIndex])

    def getHyperParameterMeanValue(self, t, name):
        """Mean of one hyper-parameter under its model's hyper-posterior at time stamp `t` (core.py:2574-2614)."""
        self._needHistory('hyper-parameter mean values', 'getCurrentHyperParameterMeanValue')
        tmIndex, hpIndex = self._locateHyperParameter(name)
        return self._hyperMean(self.hyperParameterSequence[self._timeIndex(t)][tmIndex], tmIndex, hpIndex)

    def getCurrentHyperParameterMeanValue(self, name):
        tmIndex, hpIndex = self._locateHyperParameter(name)
        return self._hyperMean(self.hyperParameterDistribution[tmIndex], tmIndex, hpIndex)

    def getHyperParameterMeanValues(self, name):
        self._needHistory('hyper-parameter mean values', 'getCurrentHyperParameterMeanValue')
        tmIndex, hpIndex = self._locateHyperParameter(name)
        return np.array([self._hyperMean(h[tmIndex], tmIndex, hpIndex) for h in self.hyperParameterSequence])

    def getHyperParameterDistribution(self, t, name, plot=False, **kwargs):
        """Marginal distribution of one hyper-parameter at time stamp `t` or 'avg' (core.py:2651-2718; like the
        reference, the stored densities are returned as they are, without the hyper-grid constant)."""
        self._needHistory('hyper-parameter distributions', 'getCurrentHyperParameterDistribution')
        tmIndex, hpIndex = self._locateHyperParameter(name)
        if isinstance(t, str) and t == 'avg':
            flat = np.mean([np.asarray(h[tmIndex], dtype=float) for h in self.hyperParameterSequence], axis=0)
        else:
            flat = self.hyperParameterSequence[self._timeIndex(t)][tmIndex]
        return self.allFlatHyperParameterValues[tmIndex][hpIndex], self._hyperMarginal(flat, tmIndex, hpIndex)

    def getHPD(self, t, name, plot=False, **kwargs):
        return self.getHyperParameterDistribution(t, name, plot=plot, **kwargs)

    def getCurrentHyperParameterDistribution(self, name, plot=False, **kwargs):
        """Marginal probabilities of one hyper-parameter after the last step (core.py:2726-2777)."""
        tmIndex, hpIndex = self._locateHyperParameter(name)
        marginal = self._hyperMarginal(self.hyperParameterDistribution[tmIndex], tmIndex, hpIndex)
        return self.allFlatHyperParameterValues[tmIndex][hpIndex], marginal * np.prod(self.hyperGridConstants[tmIndex])

    def getCHPD(self, name, plot=False, **kwargs):
        return self.getCurrentHyperParameterDistribution(name, plot=plot, **kwargs)

    def getHyperParameterDistributions(self, name):
        """Marginal probabilities of one hyper-parameter for all steps: sorted unique values and [T, #values],
        each row normalised to one (core.py:2785-2831)."""
        self._needHistory('hyper-parameter distributions', 'getCurrentHyperParameterDistribution')
        tmIndex, hpIndex = self._locateHyperParameter(name)
        column = np.asarray(self.hyperParameterValues[tmIndex], dtype=float)[:, hpIndex]
        values, inverse = np.unique(column, return_inverse=True)
        seq = np.array([np.asarray(h[tmIndex], dtype=float) for h in self.hyperParameterSequence])
        marginal = np.zeros((len(seq), len(values)))
        for k in range(len(values)):
            marginal[:, k] = np.sum(seq[:, inverse == k], axis=1)
        return values, marginal / np.sum(marginal, axis=1)[:, None]

    def getHPDs(self, name):
        return self.getHyperParameterDistributions(name)

    def getJointHyperParameterDistribution(self, names, plot=False, figure=None, subplot=111, **kwargs):
        raise NotImplementedError('This method is not available in "OnlineStudy".')  # core.py:2897-2898


def _rows_separable(ops, rows):
    """The promise behind BLG_F_SEPARABLE_ROWS (include/blgrid.h) for the step index -1 that OnlineStudy hands to
    its models: in every row the active operators are GaussianRandomWalks on distinct axes, optionally followed by
    one RegimeSwitch, or a single reset (Independent).  Such a row needs one pass over its cells and two sums, which
    is what lets the engine tile a step over the whole GPU."""
    for r in range(rows):
        active = []
        for op in ops:
            lo, hi = int(op['window'][r][0]), int(op['window'][r][1])
            if not (lo <= -1 < hi):
                continue
            if op['kind'] == _tm.OP_GRW and not (op['param'][r] > 0 and op['radius'][r] > 0):
                continue  # a random walk of zero width is the identity (transitionModels.py:110-113)
            active.append((op['kind'], op['axis']))
        kinds = [k for k, _ in active]
        if kinds == [_tm.OP_RESET]:
            continue
        if kinds and kinds[-1] == _tm.OP_REGIME:
            active = active[:-1]
        axes = [ax for k, ax in active if k == _tm.OP_GRW]
        if len(axes) != len(active) or len(set(axes)) != len(axes):
            return False
    return True


class _ResetShim:
    """Minimal stand-in so OnlineStudy can reuse _Session.reset_base without opening a full session."""

    def __init__(self, study, eng):
        self._study, self.eng, self.resetBase = study, eng, None


def _logsumexp(x):
    x = np.asarray(x, dtype=float)
    top = np.amax(x)
    if not np.isfinite(top):
        return top
    return top + np.log(np.sum(np.exp(x - top)))


def _guard_plot_arguments():
    """The accessors keep the reference's signatures, including `plot=`; plotting (matplotlib) is outside this engine,
    so `plot=True` raises instead of being silently ignored."""
    import functools
    import inspect

    def guard(fn, position):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            wanted = kwargs.get('plot', args[position] if len(args) > position else False)
            if wanted:
                raise NotImplementedError('{}(plot=True): plotting is not part of bayesloop_b200; the method returns '
                                          'the arrays the reference would plot.'.format(fn.__name__))
            return fn(*args, **kwargs)
        return wrapper

    for cls in (Study, HyperStudy, ChangepointStudy, OnlineStudy):
        for name, fn in list(vars(cls).items()):
            if inspect.isfunction(fn):
                params = list(inspect.signature(fn).parameters)
                if 'plot' in params:
                    setattr(cls, name, guard(fn, params.index('plot')))


_guard_plot_arguments()
