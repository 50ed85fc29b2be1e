"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Differential probe: odd configurations and error
paths are run through the unmodified reference and through the product (host logic -> C ABI -> CPU oracle) and the
outcomes are compared -- same exception type, or same log-evidence / means / grids / distributions to 1e-8.

    python oracle/differential_probe.py      # prints one line per probe: same | DIFF (reference vs product)

Known, deliberate differences (documented in tests/test_host_logic.py and DESIGN.md): Poisson counts > 170 overflow the
reference's factorial (OverflowError) where the product evaluates lgamma; an unknown `target` raises
ConfigurationError instead of a bare ValueError; timestamps of the wrong length fall back to the integer range (the
reference leaves them unset and crashes later); an ndarray prior is never written to by the filter (the reference's
forward loop multiplies the likelihood into the user's array, so resets and later combinations of a sweep start from
whatever it holds by then: those probes compare with the reference given the same prior as a callable).
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings('ignore')

import ref_shim  # noqa: E402


def main():
    import sympy.stats as stats
    ref = ref_shim.import_reference()
    from bayesloop_b200 import engine
    engine.set_default_engine(engine.Engine(os.path.join(HERE, 'libblgrid_oracle.so'), 'cpu'))
    import bayesloop_b200 as ours

    def run(bl, build, attrs):
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
            try:
                S = build(bl)
                return ('ok',) + tuple(np.asarray(a(S), dtype=float) for a in attrs)
            except Exception as e:  # noqa: BLE001 -- the exception type is the thing compared
                return ('exc', type(e).__name__, str(e)[:90])

    def generic(cls, data, om, tm, ts=None, **kw):
        def build(bl):
            S = getattr(bl, cls)()
            if ts is None:
                S.loadData(np.array(data))
            else:
                S.loadData(np.array(data), timestamps=ts)
            S.set(om(bl), tm(bl))
            S.fit(**kw)
            return S
        return build

    def online(data, models, prior=None, store=True):
        def build(bl):
            S = bl.OnlineStudy(storeHistory=store)
            S.setOM(G2(bl))
            for n, m in models(bl):
                S.add(n, m)
            if prior is not None:
                S.setTransitionModelPrior(prior)
            for d in data:
                S.step(d)
            return S
        return build

    def P(bl):
        return bl.om.Poisson('r', bl.oint(0, 6, 40))

    def G2(bl):
        return bl.om.Gaussian('m', bl.cint(-2, 2, 16), 's', bl.oint(0, 2, 14))

    def static(bl):
        return bl.tm.Static()

    def grw(bl, v=0.2):
        return bl.tm.GaussianRandomWalk('s', v, target='r')

    std = [lambda S: S.logEvidence, lambda S: S.posteriorMeanValues,
           lambda S: np.concatenate([np.ravel(g) for g in S.marginalGrid])]
    hyp = std + [lambda S: S.hyperParameterDistribution]
    rng = np.random.default_rng(5)
    xs = rng.normal(0.3, 0.8, 15)
    xn = xs.copy()
    xn[4] = np.nan
    xc = rng.poisson(3, 15)
    xb = rng.integers(0, 2, 15)
    xg = np.stack([xs, 0.5 + rng.random(15)], axis=1)
    ar = lambda bl: bl.om.ScaledAR1('rho', bl.oint(-1, 1, 20), 'sig', bl.oint(0, 2, 20))  # noqa: E731
    serial2 = lambda bl: bl.tm.SerialTransitionModel(  # noqa: E731
        bl.tm.Static(), bl.tm.ChangePoint('t1', 'all'), bl.tm.Static(), bl.tm.ChangePoint('t2', 'all'), bl.tm.Static())
    nested = lambda bl: bl.tm.SerialTransitionModel(  # noqa: E731
        bl.tm.GaussianRandomWalk('a', 0.2, target='r'), bl.tm.ChangePoint('t', 6),
        bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('b', 0.4, target='r'), bl.tm.RegimeSwitch('p', -5)),
        bl.tm.BreakPoint('u', 11), bl.tm.Independent())
    probes = {
        'single data point': (generic('Study', [3.], P, grw), std),
        'two points, segment of two (T = 1)': (generic('Study', [0.3, -0.2], ar, static), std),
        'one point, segment of two (T = 0)': (generic('Study', [0.3], ar, static), std),
        'all data missing': (generic('Study', [np.nan] * 3, P, grw), std),
        'grid of one cell': (generic('Study', [1., 2., 3.], lambda bl: bl.om.Poisson('r', bl.oint(0, 6, 1)), static), std),
        'grid of two cells, wide walk': (generic('Study', [1., 2., 3.], lambda bl: bl.om.Poisson('r', bl.oint(0, 6, 2)),
                                                 lambda bl: grw(bl, 5.0)), std),
        'walk far wider than the grid': (generic('Study', [1., 2., 3.], P, lambda bl: grw(bl, 100.0)), std),
        'negative Poisson count': (generic('Study', [1., -2., 3.], P, static), std),
        'change-point at the last step': (generic('Study', [1., 2., 3., 4.], P, lambda bl: bl.tm.ChangePoint('t', 3)), std),
        'change-point outside the data': (generic('Study', [1., 2., 3., 4.], P, lambda bl: bl.tm.ChangePoint('t', 17)), std),
        'hyper: one-element list': (generic('HyperStudy', [1., 2., 3., 4.], P, lambda bl: grw(bl, [0.3])), std),
        'hyper: evidenceOnly': (generic('HyperStudy', [1., 2., 3., 4.], P, lambda bl: grw(bl, [0.1, 0.3]), evidenceOnly=True),
                                [lambda S: S.logEvidence, lambda S: S.hyperParameterDistribution]),
        'hyper: forwardOnly, two hyper-parameters': (generic(
            'HyperStudy', [1, 2, 3, 2, 5, 1], P,
            lambda bl: bl.tm.CombinedTransitionModel(grw(bl, [0.1, 0.2, 0.3]), bl.tm.RegimeSwitch('p', [-7, -4])),
            forwardOnly=True), hyp + [lambda S: S.localEvidence]),
        'hyper: identical change-/break-points': (generic(
            'HyperStudy', [1, 2, 1, 6, 7, 6], P,
            lambda bl: bl.tm.SerialTransitionModel(bl.tm.Static(), bl.tm.ChangePoint('t1', [2, 3]), bl.tm.Static(),
                                                   bl.tm.ChangePoint('t2', [3, 4]), bl.tm.Static())), []),
        'changepoint study: all, 4 points': (generic('ChangepointStudy', [1., 2., 5., 6.],
                                                     lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 50)),
                                                     lambda bl: bl.tm.ChangePoint('t', 'all')), hyp),
        'changepoint study: two change-points, all': (generic('ChangepointStudy', [1, 2, 1, 6, 7, 6, 2, 1], P, serial2), hyp),
        'serial: break-points on time stamps': (generic(
            'HyperStudy', [1, 2, 1, 6, 7, 6, 2, 1], P,
            lambda bl: bl.tm.SerialTransitionModel(grw(bl, [0.1, 0.4]), bl.tm.BreakPoint('b', [1903, 1905]), bl.tm.Static()),
            ts=[1900 + i for i in range(8)]), hyp),
        'serial: nested change-point, break-point, independent': (generic(
            'Study', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60)), nested), std),
        'random walk without target': (generic('Study', [1, 2, 3], P, lambda bl: bl.tm.GaussianRandomWalk('s', 0.1)), []),
        'prior array of the wrong shape': (generic('Study', [1, 2, 3],
                                                   lambda bl: bl.om.Poisson('r', bl.oint(0, 6, 40), prior=np.ones(7)), static), []),
        'two data columns multiplied': (generic('Study', [[1, 2], [2, 3], [3, 1]], P, static), std),
        'NotEqual on a 2-D grid': (generic('Study', xs, G2, lambda bl: bl.tm.NotEqual('q', -4)), std),
        'online: missing data point': (online(xn, lambda bl: [('a', bl.tm.GaussianRandomWalk('s', [0.1, 0.2], target='m')),
                                                               ('b', bl.tm.Static())]),
                                       [lambda S: S.logEvidence, lambda S: S.marginalizedPosterior,
                                        lambda S: S.transitionModelDistribution]),
        'online: unnormalised model prior': (online(xs, lambda bl: [('a', bl.tm.RegimeSwitch('p', [-5, -3])),
                                                                     ('b', bl.tm.Independent())], prior=[2., 1.]),
                                             [lambda S: S.logEvidence, lambda S: S.transitionModelDistribution,
                                              lambda S: S.localTransitionModelDistribution]),
        'online: model prior of the wrong length': (online(xs, lambda bl: [('a', bl.tm.Static())], prior=[0.5, 0.5]), []),
        'online: duplicate hyper-parameter names': (online(
            xs, lambda bl: [('a', bl.tm.GaussianRandomWalk('s', 0.1, target='m')),
                            ('b', bl.tm.GaussianRandomWalk('s', 0.2, target='s'))]), []),
        'default grid + Jeffreys prior: Poisson': (generic('Study', xc, lambda bl: bl.om.Poisson('r'), static), std),
        'default grids: Gaussian': (generic('Study', xs, lambda bl: bl.om.Gaussian('m', None, 's', None), static), std),
        'default grids: Laplace': (generic('Study', xs, lambda bl: bl.om.Laplace('m', None, 'b', None), static), std),
        'default grid: Bernoulli': (generic('Study', xb, lambda bl: bl.om.Bernoulli('p'), static), std),
        'default grid: WhiteNoise': (generic('Study', xs, lambda bl: bl.om.WhiteNoise('s'), static), std),
        'default grid: GaussianMean': (generic('Study', xg, lambda bl: bl.om.GaussianMean('m'), static), std),
        'default grids: AR1': (generic('Study', xs, lambda bl: bl.om.AR1('rho', None, 'sig', None), static), std),
        'default grids: ScaledAR1': (generic('Study', xs, lambda bl: bl.om.ScaledAR1('rho', None, 'sig', None), static), std),
        'grid given as a number of points': (generic('Study', xc, lambda bl: bl.om.Poisson('r', 50), static), std),
        'list of SymPy priors': (generic('Study', xs, lambda bl: bl.om.Gaussian(
            'm', bl.cint(-2, 2, 20), 's', bl.oint(0, 2, 20),
            prior=[stats.Normal('a', 0, 1), stats.Exponential('b', 1)]), static), std),
    }
    # ---- second batch: fit modes, time stamps, priors / hyper-priors of every kind, irregular grids, dead combinations
    seq = std + [lambda S: S.posteriorSequence, lambda S: S.localEvidence]
    P60 = lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60))  # noqa: E731
    tsf = [0.5 * i for i in range(15)]
    tsn = [0., 1., 2.5, 2.75, 4., 7., 8., 8.5, 9., 12., 13., 14., 20., 21., 22.]

    def refit(bl):  # fit, exchange the transition model, fit again on the same study object
        S = bl.Study()
        S.loadData(np.array(xc))
        S.set(P60(bl), bl.tm.Static())
        S.fit()
        S.setTM(bl.tm.GaussianRandomWalk('s', 0.3, target='r'))
        S.fit()
        return S

    def online_hp(bl):
        S = bl.OnlineStudy(storeHistory=False)
        S.setOM(G2(bl))
        S.add('w', bl.tm.GaussianRandomWalk('s', [0.05, 0.1, 0.2], target='m', prior=lambda s: 1. / s))
        S.add('r', bl.tm.RegimeSwitch('p', [-6, -3], prior=np.array([1., 3.])))
        for d in xs:
            S.step(d)
        return S

    probes.update({
        'study: forwardOnly': (generic('Study', xc, P60, lambda bl: grw(bl, 0.3), forwardOnly=True), seq),
        'study: evidenceOnly': (generic('Study', xc, P60, lambda bl: grw(bl, 0.3), evidenceOnly=True),
                                [lambda S: S.logEvidence, lambda S: S.localEvidence]),
        'study: fit, new transition model, fit again': (refit, seq),
        'study: integer data': (generic('Study', np.asarray(xc, dtype=np.int64), P60, lambda bl: grw(bl, 0.2)), seq),
        'change-point on a fractional time stamp': (generic('Study', xc, P60, lambda bl: bl.tm.ChangePoint('t', 3.5), ts=tsf), seq),
        'change-point between two time stamps (never fires)': (generic('Study', xc, P60, lambda bl: bl.tm.ChangePoint('t', 3.3),
                                                                        ts=tsf), seq),
        'break-point between irregular time stamps': (generic(
            'Study', xc, P60, lambda bl: bl.tm.SerialTransitionModel(grw(bl, 0.5), bl.tm.BreakPoint('b', 7.5), bl.tm.Static()),
            ts=tsn), seq),
        'changepoint study: all, irregular time stamps': (generic('ChangepointStudy', xc, P60,
                                                                  lambda bl: bl.tm.ChangePoint('t', 'all'), ts=tsn), hyp),
        'random walk of width zero': (generic('Study', xc, P60, lambda bl: grw(bl, 0.0)), seq),
        'random walk of negative width': (generic('Study', xc, P60, lambda bl: grw(bl, -0.2)), seq),
        'two random walks on the same parameter': (generic(
            'Study', xc, P60, lambda bl: bl.tm.CombinedTransitionModel(grw(bl, 0.2), bl.tm.GaussianRandomWalk('s2', 0.4, target='r'))),
            seq),
        'regime switch with a floor of 1e-300': (generic('Study', xc, P60, lambda bl: bl.tm.RegimeSwitch('p', -300)), seq),
        'regime switch with a floor above the flat level': (generic('Study', xc, P60, lambda bl: bl.tm.RegimeSwitch('p', 1)), seq),
        'independent observations': (generic('Study', xc, P60, lambda bl: bl.tm.Independent()), seq),
        'static inside a combined model': (generic(
            'Study', xc, P60, lambda bl: bl.tm.CombinedTransitionModel(bl.tm.Static(), grw(bl, 0.3), bl.tm.Static())), seq),
        'irregular (categorical) grid with a random walk': (generic(
            'Study', xc, lambda bl: bl.om.Poisson('r', [0.5, 1., 2., 3., 3.5, 5., 8.]), lambda bl: grw(bl, 1.2)), seq),
        '2-D grid, one irregular axis': (generic(
            'Study', xs, lambda bl: bl.om.Gaussian('m', bl.cint(-2, 2, 12), 's', [0.2, 0.5, 0.6, 1., 2.]),
            lambda bl: bl.tm.GaussianRandomWalk('w', 0.3, target='m')), seq),
        'callable prior on a 2-D grid': (generic('Study', xs, lambda bl: bl.om.Gaussian(
            'm', bl.cint(-2, 2, 16), 's', bl.oint(0, 2, 14), prior=lambda m, s: 1. / s ** 3), static), seq),
        'SymPy prior, one parameter': (generic('Study', xc, lambda bl: bl.om.Poisson(
            'r', bl.oint(0, 8, 60), prior=stats.Exponential('e', 0.5)), lambda bl: grw(bl, 0.2)), seq),
        'prior array that sums to one': (generic('Study', xc, lambda bl: bl.om.Poisson(
            'r', bl.oint(0, 8, 60), prior=np.ones(60) / 60.), static), seq),
        # the reference's forward loop works IN PLACE on the array _computePrior returned (core.py:362, :385): with an
        # ndarray prior that is the user's array itself, so a later reset (transitionModels.py:303-304, :352-353)
        # "restores" whatever the filter has written into it by then.  The product keeps the prior the user gave;
        # the reference run it is compared with gets the same values through a callable (core.py:224-235), which the
        # reference evaluates afresh at every reset
        'prior array with a zero region and a change-point': (
            generic('Study', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=np.concatenate([np.zeros(10), np.ones(50)])),
                    lambda bl: bl.tm.ChangePoint('t', 6)), seq,
            generic('Study', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=lambda r: 1. * (np.arange(60) >= 10)),
                    lambda bl: bl.tm.ChangePoint('t', 6))),
        'prior array and independent observations': (
            generic('Study', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=np.linspace(1., 3., 60)),
                    lambda bl: bl.tm.Independent()), seq,
            generic('Study', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=lambda r: np.linspace(1., 3., 60)),
                    lambda bl: bl.tm.Independent())),
        'hyper: prior array (every combination starts from it)': (
            generic('HyperStudy', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=np.linspace(1., 3., 60)),
                    lambda bl: grw(bl, [0.1, 0.3])), hyp + [lambda S: S.posteriorSequence],
            generic('HyperStudy', xc, lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 60), prior=lambda r: np.linspace(1., 3., 60)),
                    lambda bl: grw(bl, [0.1, 0.3]))),
        'count the likelihood cannot explain (dead fit)': (generic(
            'Study', [1., 2., 900., 1.], lambda bl: bl.om.Gaussian('m', bl.cint(-2, 2, 16), 's', bl.oint(0, 0.5, 10)), static),
            [lambda S: S.logEvidence]),
        'hyper: one dead combination of three': (generic(
            'HyperStudy', [0.1, 0.2, 30., 0.1], lambda bl: bl.om.Gaussian('m', bl.cint(-1, 1, 16), 's', bl.oint(0, 0.4, 8)),
            lambda bl: bl.tm.GaussianRandomWalk('w', [0., 0.01, 4.], target='m')),
            [lambda S: S.logEvidence, lambda S: S.hyperParameterDistribution, lambda S: np.asarray(S.logEvidenceList)]),
        'hyper: callable hyper-prior': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.GaussianRandomWalk(
            's', [0.1, 0.2, 0.4], target='r', prior=lambda s: 1. / s)), hyp + [lambda S: S.posteriorSequence]),
        'hyper: hyper-prior array': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.GaussianRandomWalk(
            's', [0.1, 0.2, 0.4], target='r', prior=np.array([3., 2., 1.]))), hyp),
        'hyper: SymPy hyper-prior': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.GaussianRandomWalk(
            's', bl.cint(0.1, 0.5, 5), target='r', prior=stats.Exponential('e', 4.))), hyp),
        'hyper: irregular hyper-parameter values': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.GaussianRandomWalk(
            's', [0.1, 0.15, 0.4, 1.0], target='r')), hyp + [lambda S: S.localEvidence]),
        'hyper: break-point all inside a hyper-study': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.SerialTransitionModel(
            bl.tm.Static(), bl.tm.BreakPoint('b', 'all'), grw(bl, [0.1, 0.3]))), hyp),
        'changepoint study: hyper-prior on the change-point': (generic(
            'ChangepointStudy', xc, P60, lambda bl: bl.tm.ChangePoint('t', 'all', prior=lambda t: 1. + t)), hyp),
        'changepoint study: change-point x walk widths (shared prefixes)': (generic(
            'ChangepointStudy', xc, P60, lambda bl: bl.tm.CombinedTransitionModel(
                bl.tm.ChangePoint('t', 'all'), grw(bl, [0.1, 0.2, 0.4]))), hyp + [lambda S: S.posteriorSequence,
                                                                                   lambda S: S.localEvidence]),
        'online: hyper-priors, no history': (online_hp, [lambda S: S.logEvidence, lambda S: S.marginalizedPosterior,
                                                         lambda S: S.transitionModelDistribution,
                                                         lambda S: np.concatenate(S.hyperParameterDistribution)]),
        'online: segment of two, first point only waits': (
            lambda bl: [S for S in [bl.OnlineStudy()] if not (S.setOM(ar(bl)), S.add('a', bl.tm.Static()), S.step(0.3))][0],
            [lambda S: float(len(S.formattedTimestamps))]),
        'missing value in a segment of two': (generic('Study', xn, ar, static), seq),
        'Bernoulli with a value outside {0, 1}': (generic('Study', [0, 1, 2, 1], lambda bl: bl.om.Bernoulli('p', bl.oint(0, 1, 20)),
                                                          static), std),
    })
    # ---- third batch: OnlineStudy with history and odd models, optimize / simulate, accessors, wider sweeps
    def online_full(models, data=xs, om=G2, prior=None):
        def build(bl):
            S = bl.OnlineStudy(storeHistory=True)
            S.setOM(om(bl))
            for n, m in models(bl):
                S.add(n, m)
            if prior is not None:
                S.setTransitionModelPrior(prior)
            for d in data:
                S.step(d)
            return S
        return build

    hist = [lambda S: S.logEvidence, lambda S: np.asarray(S.posteriorSequence), lambda S: np.asarray(S.posteriorMeanValues),
            lambda S: np.asarray(S.transitionModelSequence), lambda S: np.asarray(S.localTransitionModelSequence),
            lambda S: np.concatenate([np.concatenate(h) for h in S.hyperParameterSequence])]
    two = lambda bl: [('w', bl.tm.CombinedTransitionModel(  # noqa: E731
        bl.tm.GaussianRandomWalk('a', [0.05, 0.2], target='m'), bl.tm.GaussianRandomWalk('b', [0.02, 0.1, 0.3], target='s'))),
        ('r', bl.tm.RegimeSwitch('p', [-6, -2])), ('i', bl.tm.Independent()), ('n', bl.tm.NotEqual('q', -3.))]  # (an integer exponent makes the reference's 10**q raise)

    def online_set(bl):  # transition model given through set(), not add()
        S = bl.OnlineStudy(storeHistory=True)
        S.set(G2(bl), bl.tm.GaussianRandomWalk('s', 0.1, target='m'))
        for d in xs:
            S.step(d)
        return S

    def optimized(bl):
        S = bl.Study()
        S.loadData(np.array(xc))
        S.set(P60(bl), bl.tm.GaussianRandomWalk('s', 0.5, target='r'))
        S.optimize(tol=1e-6)
        return S

    def custom_grid(bl):
        S = bl.HyperStudy()
        S.loadData(np.array(xc))
        S.set(P60(bl), bl.tm.CombinedTransitionModel(grw(bl, [0.1, 0.2]), bl.tm.RegimeSwitch('p', [-7, -4])))
        S._createHyperGrid()
        S.hyperGridValues = np.array([[0.1, -7.], [0.15, -5.], [0.3, -4.]])
        S.flatHyperPriorValues = np.array([0.2, 0.5, 0.3])
        S.hyperGridConstant = [1., 1.]
        S.fit(customHyperGrid=True)
        return S

    fitted = lambda bl: generic('HyperStudy', xc, P60, lambda b: grw(b, [0.1, 0.2, 0.4]))(bl)  # noqa: E731
    fitted1 = lambda bl: generic('Study', xc, P60, lambda b: grw(b, 0.3))(bl)  # noqa: E731
    cps2 = lambda bl: generic('ChangepointStudy', [1, 2, 1, 6, 7, 6, 2, 1, 1, 2], P, lambda b: b.tm.SerialTransitionModel(  # noqa: E731
        b.tm.Static(), b.tm.ChangePoint('t1', 'all'), grw(b, [0.1, 0.3]), b.tm.ChangePoint('t2', 'all'), b.tm.Static()))(bl)
    probes.update({
        'online: history of four hypotheses families': (online_full(two), hist),
        'online: model prior and history': (online_full(two, prior=[0.1, 0.2, 0.3, 0.4]), hist),
        'online: transition model through set()': (online_set, hist[:3]),
        'online: change-point at time -1 (t is always -1)': (online_full(
            lambda bl: [('c', bl.tm.ChangePoint('t', -1)), ('s', bl.tm.Static())]), hist[:5]),
        'online: serial model (t is always -1)': (online_full(
            lambda bl: [('x', bl.tm.SerialTransitionModel(bl.tm.Static(), bl.tm.BreakPoint('b', 5),
                                                          bl.tm.GaussianRandomWalk('s', 0.2, target='m')))]), hist[:3]),
        'online: Poisson, one parameter': (online_full(
            lambda bl: [('w', bl.tm.GaussianRandomWalk('s', [0.1, 0.3], target='r')), ('r', bl.tm.RegimeSwitch('p', -5))],
            data=xc, om=P60), hist),
        'online: segment of two': (online_full(
            lambda bl: [('w', bl.tm.GaussianRandomWalk('s', [0.05, 0.1], target='rho')), ('s', bl.tm.Static())], om=ar), hist),
        'online: two data columns per step': (online_full(
            lambda bl: [('w', bl.tm.GaussianRandomWalk('s', [0.1, 0.3], target='r'))],
            data=[[1, 2], [2, 3], [3, 1], [0, 2]], om=P60), hist[:3]),
        'optimize one hyper-parameter': (optimized, [lambda S: S.logEvidence, lambda S: S.getHyperParameterValue('s'),
                                                     lambda S: S.posteriorMeanValues]),
        'hyper: custom hyper-grid': (custom_grid, hyp + [lambda S: S.posteriorSequence]),
        'hyper: three hyper-parameters': (generic('HyperStudy', xs, G2, lambda bl: bl.tm.CombinedTransitionModel(
            bl.tm.GaussianRandomWalk('a', [0.05, 0.2], target='m'), bl.tm.GaussianRandomWalk('b', [0.02, 0.1], target='s'),
            bl.tm.RegimeSwitch('p', [-6, -3, -1]))), hyp + [lambda S: S.posteriorSequence, lambda S: S.localEvidence]),
        'hyper: NotEqual sweep': (generic('HyperStudy', xs, G2, lambda bl: bl.tm.NotEqual('q', [-5., -3., -1.])), hyp),
        'hyper: every combination dead': (generic(
            'HyperStudy', [0.1, 50., 0.1], lambda bl: bl.om.Gaussian('m', bl.cint(-1, 1, 16), 's', bl.oint(0, 0.3, 8)),
            lambda bl: bl.tm.GaussianRandomWalk('w', [0., 0.01], target='m')), [lambda S: S.logEvidence]),
        'hyper: change-point values in a plain hyper-study': (generic('HyperStudy', xc, P60, lambda bl: bl.tm.CombinedTransitionModel(
            bl.tm.ChangePoint('t', [3, 6, 9]), grw(bl, [0.1, 0.3]))), hyp + [lambda S: S.posteriorSequence]),
        'changepoint study: evidenceOnly': (generic('ChangepointStudy', xc, P60, lambda bl: bl.tm.ChangePoint('t', 'all'),
                                                    evidenceOnly=True), [lambda S: S.logEvidence, lambda S: S.hyperParameterDistribution]),
        'changepoint study: two change-points around a random walk': (cps2, hyp + [lambda S: S.posteriorSequence]),
        'accessor: getParameterMeanValues': (fitted, [lambda S: S.getParameterMeanValues('r')]),
        # (getParameterDistribution(s) of the reference compare an ndarray with [] and raise under NumPy 2: core.py:880, :950)
        'accessor: getHyperParameterDistribution': (fitted, [lambda S: S.getHyperParameterDistribution('s')[0],
                                                             lambda S: S.getHyperParameterDistribution('s')[1]]),
        'accessor: getJointHyperParameterDistribution': (
            generic('HyperStudy', xc, P60, lambda bl: bl.tm.CombinedTransitionModel(grw(bl, [0.1, 0.2, 0.3]),
                                                                                   bl.tm.RegimeSwitch('p', [-7, -4]))),
            [lambda S: S.getJointHyperParameterDistribution(['s', 'p'])[2]]),
        'accessor: getDurationDistribution': (cps2, [lambda S: S.getDurationDistribution(['t1', 't2'])[0],
                                                      lambda S: S.getDurationDistribution(['t1', 't2'])[1]]),
        'accessor: unknown parameter name': (fitted, [lambda S: S.getParameterMeanValues('nope')]),
        'accessor: online current distributions': (online_full(two), [
            lambda S: S.getCurrentParameterDistribution('m')[1], lambda S: S.getCurrentTransitionModelDistribution(),
            lambda S: S.getCurrentTransitionModelProbability('r'), lambda S: S.getCurrentParameterMeanValue('s'),
            lambda S: S.getCurrentHyperParameterDistribution('a')[1], lambda S: S.getCurrentHyperParameterMeanValue('b'),
            lambda S: S.getTransitionModelProbabilities('w'), lambda S: S.getParameterMeanValues('m'),
            lambda S: S.getHyperParameterMeanValues('p'), lambda S: S.getParameterDistribution(3, 's')[1],
            lambda S: S.getHyperParameterDistribution(5, 'b')[1]]),
        'simulate: Poisson predictions': (fitted1, [lambda S: np.asarray(S.simulate(np.arange(0, 12), t=5, density=False))[1]]),
    })
    diffs = 0
    for name, probe in probes.items():
        build, attrs = probe[:2]
        r, o = run(ref, probe[2] if len(probe) > 2 else build, attrs), run(ours, build, attrs)
        if r[0] != o[0]:
            same = False
        elif r[0] == 'exc':
            same = r[1] == o[1]
        else:
            same = all(x.shape == y.shape and np.allclose(x, y, rtol=1e-8, atol=1e-300 if x.ndim else 1e-12, equal_nan=True)
                       for x, y in zip(r[1:], o[1:]))
        diffs += not same
        detail = '' if same else '   reference: %s | product: %s' % (r[:3] if r[0] == 'exc' else 'ok', o[:3] if o[0] == 'exc' else 'ok')
        print('%-52s %s%s' % (name, 'same' if same else 'DIFF', detail))
    print('%d probes, %d differ' % (len(probes), diffs))


if __name__ == '__main__':
    main()
