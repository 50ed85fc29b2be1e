"""Random study configurations for the differential tests: a study type, an observation model, a random tree of
transition models (Combined / Serial with change- and break-points, nested, hyper-parameter lists, 'all' change-points),
data with optional gaps and time stamps.  `draw_case(seed)` returns a builder `bl -> fitted study` that works with the
reference package and with the product alike (used by oracle/fuzz_lowering.py against the unmodified reference and by
tests/test_gpu_fuzz.py for CUDA against the CPU oracle).  No dependency on the reference."""
import numpy as np


class Namer:
    def __init__(self):
        self.k = 0

    def __call__(self, stem):
        self.k += 1
        return '%s%d' % (stem, self.k)


def draw_leaf(rng, names, params, hyper, T, t0):
    """Returns a builder bl -> model; hyper=True draws lists of values."""
    kind = rng.choice(['static', 'grw', 'grw', 'regime', 'cp', 'indep', 'notequal'])
    if kind == 'static':
        return lambda bl: bl.tm.Static()
    if kind == 'grw':
        name, target = names('s'), str(rng.choice(params))
        vals = sorted(set(np.round(rng.uniform(0.0, 0.6, rng.integers(2, 4)), 3))) if hyper else float(np.round(rng.uniform(0, 0.6), 3))
        return lambda bl: bl.tm.GaussianRandomWalk(name, vals, target=target)
    if kind == 'regime':
        name = names('p')
        vals = [-7., -4.] if hyper and rng.random() < 0.5 else -5.
        return lambda bl: bl.tm.RegimeSwitch(name, vals)
    if kind == 'cp':
        name = names('t')
        if hyper and rng.random() < 0.6:
            vals = sorted(set(int(v) for v in rng.integers(t0, t0 + T, 3)))
        else:
            vals = int(rng.integers(t0, t0 + T))
        return lambda bl: bl.tm.ChangePoint(name, vals)
    if kind == 'indep':
        return lambda bl: bl.tm.Independent()
    name = names('q')
    return lambda bl: bl.tm.NotEqual(name, -4.)


def draw_tree(rng, names, params, hyper, T, t0, depth=0):
    r = rng.random()
    if depth >= 2 or r < 0.35:
        return draw_leaf(rng, names, params, hyper, T, t0)
    if r < 0.7:
        subs = [draw_tree(rng, names, params, hyper, T, t0, depth + 1) for _ in range(rng.integers(2, 4))]
        return lambda bl: bl.tm.CombinedTransitionModel(*[s(bl) for s in subs])
    n = int(rng.integers(2, 4))
    subs = [draw_tree(rng, names, params, hyper, T, t0, depth + 1) for _ in range(n)]
    points = sorted(set(int(v) for v in rng.integers(t0 + 1, t0 + T - 1, n - 1)))
    while len(points) < n - 1:
        points = sorted(set(points + [int(rng.integers(t0 + 1, t0 + T - 1))]))
    kinds = [rng.random() < 0.5 for _ in points]
    pnames = [names('c' if k else 'b') for k in kinds]

    def build(bl):
        args = []
        for i, s in enumerate(subs):
            args.append(s(bl))
            if i < n - 1:
                cls = bl.tm.ChangePoint if kinds[i] else bl.tm.BreakPoint
                args.append(cls(pnames[i], points[i]))
        return bl.tm.SerialTransitionModel(*args)
    return build


def draw_changepoint_tree(rng, names, params, T, t0):
    """Serial model with one or two change-/break-points whose values are 'all' or explicit lists (the ordered-tuple
    mask of ChangepointStudy.fit, core.py:1765-1852), hyper-parameter lists inside the segments."""
    n = int(rng.integers(2, 4))
    subs = [draw_tree(rng, names, params, True, T, t0, depth=1) if rng.random() < 0.5 else
            (lambda bl: bl.tm.Static()) for _ in range(n)]
    subs = [s for s in subs]
    kinds = [rng.random() < 0.5 for _ in range(n - 1)]
    pnames = [names('c' if k else 'b') for k in kinds]
    values = ['all' if rng.random() < 0.6 else sorted(set(int(v) for v in rng.integers(t0, t0 + T, 4))) for _ in range(n - 1)]

    def build(bl):
        args = []
        for i, s in enumerate(subs):
            args.append(s(bl))
            if i < n - 1:
                cls = bl.tm.ChangePoint if kinds[i] else bl.tm.BreakPoint
                args.append(cls(pnames[i], values[i]))
        return bl.tm.SerialTransitionModel(*args)
    return build


def draw_case(seed):
    rng = np.random.default_rng(seed)
    names = Namer()
    study = str(rng.choice(['Study', 'Study', 'HyperStudy', 'HyperStudy', 'OnlineStudy', 'ChangepointStudy']))
    T = int(rng.integers(6, 14))
    t0 = int(rng.choice([0, 0, 1900]))
    which = str(rng.choice(['poisson', 'gauss', 'ar1', 'bernoulli']))
    if which == 'poisson':
        data = rng.poisson(3, T).astype(float)
        params = ['r']
        om = lambda bl: bl.om.Poisson('r', bl.oint(0, 8, int(rng_grid[0])))  # noqa: E731
    elif which == 'gauss':
        data = rng.normal(0.2, 0.9, T)
        params = ['m', 's']
        om = lambda bl: bl.om.Gaussian('m', bl.cint(-3, 3, int(rng_grid[0])), 's', bl.oint(0, 3, int(rng_grid[1])))  # noqa: E731
    elif which == 'ar1':
        data = rng.normal(0, 1, T + 1)
        params = ['rho', 'sig']
        om = lambda bl: bl.om.ScaledAR1('rho', bl.oint(-1, 1, int(rng_grid[0])), 'sig', bl.oint(0, 3, int(rng_grid[1])))  # noqa: E731
    else:
        data = rng.integers(0, 2, T).astype(float)
        params = ['p']
        om = lambda bl: bl.om.Bernoulli('p', bl.oint(0, 1, int(rng_grid[0])))  # noqa: E731
    rng_grid = rng.integers(8, 22, 2)
    if rng.random() < 0.3 and which != 'bernoulli':
        data[int(rng.integers(1, T - 1))] = np.nan
    hyper = study == 'HyperStudy'
    online = study == 'OnlineStudy'
    stamps = None if (t0 == 0 or online) else np.arange(t0, t0 + len(data))
    if online:
        t0 = 0
        models = [draw_tree(rng, names, params, True, T, 0, depth=1) if rng.random() < 0.4 else
                  draw_leaf(rng, names, params, True, T, 0) for _ in range(int(rng.integers(1, 4)))]
    elif study == 'ChangepointStudy':
        tree = draw_changepoint_tree(rng, names, params, T, t0 + (1 if which == 'ar1' else 0))
    else:
        tree = draw_tree(rng, names, params, hyper, T, t0 + (1 if which == 'ar1' else 0))
    kw = {}
    if not online and rng.random() < 0.3:
        kw = {'forwardOnly': True} if rng.random() < 0.5 else {'evidenceOnly': True}

    def build(bl):
        if online:
            S = bl.OnlineStudy(storeHistory=True)
            S.setOM(om(bl))
            for k, m in enumerate(models):
                S.add('tm%d' % k, m(bl))
            for d in data:
                S.step(d)
            return S
        S = getattr(bl, study)()
        if stamps is None:
            S.loadData(data)
        else:
            S.loadData(data, timestamps=stamps)
        S.set(om(bl), tree(bl))
        S.fit(**kw)
        return S
    return build, '%s/%s T=%d %s' % (study, which, T, kw or '')


def extract(S):
    out = [np.asarray(S.logEvidence, dtype=float)]
    if type(S).__name__ == 'OnlineStudy':
        out += [np.asarray(S.posteriorMeanValues, dtype=float), np.asarray(S.transitionModelDistribution, dtype=float),
                np.asarray(S.marginalizedPosterior, dtype=float)]
        return out
    if np.isfinite(S.logEvidence) and isinstance(S.posteriorMeanValues, np.ndarray) and S.posteriorMeanValues.size:
        out.append(np.asarray(S.posteriorMeanValues, dtype=float))
    hp = getattr(S, 'hyperParameterDistribution', None)
    if hp is not None and len(getattr(S, 'logEvidenceList', [])) > 0:
        out.append(np.asarray(hp, dtype=float))
        out.append(np.asarray(S.logEvidenceList, dtype=float))
    return out
