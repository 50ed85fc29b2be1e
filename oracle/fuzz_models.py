"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Randomised differential test of the observation
models and priors: every built-in model (Poisson, Gaussian, Laplace, GaussianMean,
WhiteNoise, Bernoulli, AR1, ScaledAR1), the NumPy / SciPy / SymPy wrappers (likelihood-table path), default or given
grids, every kind of prior (default / None / array / callable / SymPy), one to three data columns, missing values, and a
Static / GaussianRandomWalk / RegimeSwitch / ChangePoint transition -- fitted by the unmodified reference and by the
product (host logic -> C ABI -> CPU oracle).  Log-evidence, posterior sequence, local evidence and grids must agree to
1e-8, or both sides must raise the same exception type.

    python oracle/fuzz_models.py [n_cases=300] [seed=0]

An ndarray prior is not combined with a ChangePoint: the reference's filter writes into the user's array, so its reset
restores the current posterior instead of the prior (oracle/differential_probe.py, tests/test_host_logic.py).
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

KINDS = ['poisson', 'gauss', 'laplace', 'gmean', 'wn', 'bern', 'ar1', 'sar1', 'scipy_norm', 'scipy_pois', 'sympy_norm', 'numpy']
FIRST_NAME = {'scipy_norm': 'loc', 'scipy_pois': 'mu', 'sympy_norm': 'mu'}


def draw(seed):
    import scipy.stats
    import sympy
    import sympy.stats
    rng = np.random.default_rng(seed)
    T = int(rng.integers(5, 14))
    kind = str(rng.choice(KINDS))
    n = [int(x) for x in rng.integers(5, 14, 2)]
    defaults = rng.random() < 0.2
    prior_kind = str(rng.choice(['default', 'default', 'none', 'array', 'callable', 'sympy']))
    cols = int(rng.choice([1, 1, 1, 2, 3]))
    two = kind in ('gauss', 'laplace', 'ar1', 'sar1', 'scipy_norm', 'sympy_norm')
    if kind == 'gmean':  # rows of [value, std]
        data, cols = np.stack([rng.normal(0.3, 0.9, T), 0.5 + rng.random(T)], axis=1), 1
    else:
        if kind in ('poisson', 'scipy_pois'):
            data = rng.poisson(3, (T, cols)).astype(float)
        elif kind == 'bern':
            data = rng.integers(0, 2, (T, cols)).astype(float)
        else:
            data = rng.normal(0.3, 0.9, (T, cols))
        if cols == 1 or kind in ('ar1', 'sar1'):
            data, cols = data[:, 0], 1
    if rng.random() < 0.3 and kind != 'bern':
        i = int(rng.integers(0, T))
        if data.ndim == 1:
            data[i] = np.nan
        else:
            data[i, int(rng.integers(0, data.shape[1]))] = np.nan
    transition = str(rng.choice(['static', 'grw', 'rs', 'cp']))
    if prior_kind == 'array' and transition == 'cp':
        transition = 'grw'

    def prior(shape):
        if prior_kind == 'default':
            return {}
        if prior_kind == 'none':
            return {'prior': None}
        if prior_kind == 'array':
            return {'prior': 0.5 + np.arange(np.prod(shape), dtype=float).reshape(shape) / np.prod(shape)}
        if prior_kind == 'callable':
            return {'prior': (lambda a, b: 1. / (1. + a * a) / (0.1 + b)) if two else (lambda a: 1. / (0.1 + np.abs(a)))}
        if two:
            return {'prior': [sympy.stats.Normal('pa', 0, 2), sympy.stats.Exponential('pb', 1)]}
        return {'prior': sympy.stats.Exponential('pa', 0.5)}

    def om(bl):
        grid = lambda values: None if defaults else values  # noqa: E731
        one, both = prior((n[0],)), prior((n[0], n[1]))
        if kind == 'poisson':
            return bl.om.Poisson('a', grid(bl.oint(0, 8, n[0])), **one)
        if kind == 'gauss':
            return bl.om.Gaussian('a', grid(bl.cint(-3, 3, n[0])), 'b', grid(bl.oint(0, 3, n[1])), **both)
        if kind == 'laplace':
            return bl.om.Laplace('a', grid(bl.cint(-3, 3, n[0])), 'b', grid(bl.oint(0, 3, n[1])), **both)
        if kind == 'gmean':
            return bl.om.GaussianMean('a', grid(bl.cint(-3, 3, n[0])), **one)
        if kind == 'wn':
            return bl.om.WhiteNoise('a', grid(bl.oint(0, 3, n[0])), **one)
        if kind == 'bern':
            return bl.om.Bernoulli('a', grid(bl.oint(0, 1, n[0])), **one)
        if kind == 'ar1':
            return bl.om.AR1('a', grid(bl.oint(-1, 1, n[0])), 'b', grid(bl.oint(0, 3, n[1])), **both)
        if kind == 'sar1':
            return bl.om.ScaledAR1('a', grid(bl.oint(-1, 1, n[0])), 'b', grid(bl.oint(0, 3, n[1])), **both)
        if kind == 'scipy_norm':
            return bl.om.SciPy(scipy.stats.norm, 'loc', bl.cint(-3, 3, n[0]), 'scale', bl.oint(0, 3, n[1]), **both)
        if kind == 'scipy_pois':
            return bl.om.SciPy(scipy.stats.poisson, 'mu', bl.oint(0, 8, n[0]), fixedParameters={'loc': 0}, **one)
        if kind == 'sympy_norm':
            mu, sd = sympy.Symbol('mu'), sympy.Symbol('sd', positive=True)
            return bl.om.SymPy(sympy.stats.Normal('x', mu, sd), 'mu', bl.cint(-3, 3, n[0]), 'sd', bl.oint(0, 3, n[1]), **both)
        return bl.om.NumPy(lambda x, a: np.exp(-np.abs(x - a)) / 2., 'a', bl.cint(-3, 3, n[0]), **one)

    def tm(bl):
        if transition == 'static':
            return bl.tm.Static()
        if transition == 'grw':
            return bl.tm.GaussianRandomWalk('s', 0.3, target=FIRST_NAME.get(kind, 'a'))
        if transition == 'rs':
            return bl.tm.RegimeSwitch('p', -4.)
        return bl.tm.ChangePoint('t', int(T // 2))

    def build(bl):
        S = bl.Study()
        S.loadData(np.array(data))
        S.set(om(bl), tm(bl))
        S.fit()
        return S
    return build, '%s prior=%s default grids=%s columns=%d %s T=%d' % (kind, prior_kind, defaults, cols, transition, T)


def run(fit):
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
        try:
            return 'ok', fit()
        except Exception as e:  # noqa: BLE001 -- the exception type is what is compared
            return 'exc', type(e).__name__


def same(R, O):
    if not np.isclose(R.logEvidence, O.logEvidence, rtol=1e-8, atol=1e-10, equal_nan=True):
        return False
    if not np.isfinite(R.logEvidence):
        return True
    a, b = np.asarray(R.posteriorSequence, float), np.asarray(O.posteriorSequence, float)
    grids = [np.concatenate([np.ravel(g) for g in S.marginalGrid]) for S in (R, O)]
    return (a.shape == b.shape and bool(np.all(np.abs(a - b) <= 1e-8 * np.abs(a) + 1e-12 * np.nanmax(np.abs(a))))
            and np.allclose(R.localEvidence, O.localEvidence, rtol=1e-8, equal_nan=True)
            and grids[0].shape == grids[1].shape and np.allclose(grids[0], grids[1]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ref = ref_shim.import_reference()
    from bayesloop_b200 import engine
    engine.set_default_engine(engine.Engine(os.path.join(HERE, 'libblgrid_oracle.so'), 'cpu'))
    import bayesloop_b200 as ours
    differ = rejected = 0
    for seed in range(seed0, seed0 + n):
        build, label = draw(seed)
        r, o = run(lambda: build(ref)), run(lambda: build(ours))
        if r[0] == 'exc' or o[0] == 'exc':
            rejected += r[0] == 'exc'
            if r != o:
                differ += 1
                print('seed %d: reference %s | product %s   %s' % (seed, r[1] if r[0] == 'exc' else 'ok',
                                                                     o[1] if o[0] == 'exc' else 'ok', label))
        elif not same(r[1], o[1]):
            differ += 1
            print('seed %d DIFF  %r vs %r   %s' % (seed, r[1].logEvidence, o[1].logEvidence, label))
    print('%d cases, %d rejected by the reference itself, %d differ' % (n, rejected, differ))
    return 1 if differ else 0


if __name__ == '__main__':
    sys.exit(main())
