"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Randomised differential test of OnlineStudy.step
(reference core.py:2062-2226): one to four transition models per study -- leaves or
combined models out of Static / GaussianRandomWalk / RegimeSwitch / NotEqual / Independent / ChangePoint (t is always -1
in an online study) with hyper-parameter lists and hyper-priors (none / callable / array) --, an optional
transition-model prior, with or without history, missing values, Poisson / Gaussian / ScaledAR1 grids.  The unmodified
reference and the product (host logic -> C ABI -> CPU oracle) must agree to 1e-8 on the log-evidence, the marginalised
posterior, the (local) transition-model distribution, every hyper-parameter distribution and per-hypothesis evidence,
and on the stored sequences.

    python oracle/fuzz_online.py [n_cases=300] [seed=0]

One tolerated difference: a hypothesis whose norm turns NaN (NotEqual applied to a flat distribution after a missing
value: 0/0) carries log-evidence NaN in the reference and -inf (dead) in the product; everything derived from it is NaN
on both sides.
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


def draw(seed):
    rng = np.random.default_rng(seed)
    T = int(rng.integers(4, 12))
    which = str(rng.choice(['poisson', 'gauss', 'sar1']))
    g = [int(x) for x in rng.integers(5, 12, 2)]
    if which == 'poisson':
        data, params = rng.poisson(3, T).astype(float), ['r']
        om = lambda bl: bl.om.Poisson('r', bl.oint(0, 8, g[0]))  # noqa: E731
    elif which == 'gauss':
        data, params = rng.normal(0.2, 0.9, T), ['m', 's']
        om = lambda bl: bl.om.Gaussian('m', bl.cint(-3, 3, g[0]), 's', bl.oint(0, 3, g[1]))  # noqa: E731
    else:
        data, params = rng.normal(0, 1, T + 1), ['rho', 'sig']
        om = lambda bl: bl.om.ScaledAR1('rho', bl.oint(-1, 1, g[0]), 'sig', bl.oint(0, 3, g[1]))  # noqa: E731
    if rng.random() < 0.3:
        data[int(rng.integers(1, len(data)))] = np.nan
    count = [0]

    def leaf():
        kind = str(rng.choice(['static', 'grw', 'grw', 'rs', 'ne', 'indep', 'cp']))
        count[0] += 1
        name = 'h%d' % count[0]
        prior_kind, prior_seed = str(rng.choice(['none', 'none', 'callable', 'array'])), int(rng.integers(0, 2 ** 31))
        if kind == 'grw':
            values = sorted(set(float(x) for x in np.round(rng.uniform(0, 0.5, int(rng.integers(1, 4))), 3)))
            return 'grw', name, values, str(rng.choice(params)), prior_kind, prior_seed
        if kind == 'rs':
            values = sorted(float(x) for x in rng.choice([-8., -6., -4., -2.], int(rng.integers(1, 3)), replace=False))
            return 'rs', name, values, None, prior_kind, prior_seed
        if kind == 'ne':
            return 'ne', name, -3.
        if kind == 'cp':
            return 'cp', name, int(rng.choice([-1, -1, 3]))
        return (kind,)

    specs = []
    for _ in range(int(rng.integers(1, 5))):
        specs.append(('comb', [leaf() for _ in range(int(rng.integers(2, 4)))]) if rng.random() < 0.35 else leaf())
    model_prior = [float(x) for x in 0.2 + rng.random(len(specs))] if rng.random() < 0.4 else None
    store = bool(rng.random() < 0.6)

    def make(bl, spec):
        if spec[0] == 'comb':
            return bl.tm.CombinedTransitionModel(*[make(bl, x) for x in spec[1]])
        if spec[0] == 'static':
            return bl.tm.Static()
        if spec[0] == 'indep':
            return bl.tm.Independent()
        if spec[0] == 'ne':
            return bl.tm.NotEqual(spec[1], spec[2])
        if spec[0] == 'cp':
            return bl.tm.ChangePoint(spec[1], spec[2])
        values, kw = spec[2], {}
        if len(values) > 1 and spec[4] == 'callable':
            kw = {'prior': lambda x: 1. / (1. + np.abs(x))}
        elif len(values) > 1 and spec[4] == 'array':  # normalised in place by the reference: a fresh one per build
            kw = {'prior': 0.25 + np.random.default_rng(spec[5]).random(len(values))}
        values = values if len(values) > 1 else values[0]
        if spec[0] == 'grw':
            return bl.tm.GaussianRandomWalk(spec[1], values, target=spec[3], **kw)
        return bl.tm.RegimeSwitch(spec[1], values, **kw)

    def build(bl):
        S = bl.OnlineStudy(storeHistory=store)
        S.setOM(om(bl))
        for i, spec in enumerate(specs):
            S.add('tm%d' % i, make(bl, spec))
        if model_prior is not None:
            S.setTransitionModelPrior(model_prior)
        for d in data:
            S.step(d)
        return S
    shape = [s[0] if s[0] != 'comb' else [x[0] for x in s[1]] for s in specs]
    return build, '%s T=%d history=%s model prior=%s %s' % (which, T, store, model_prior is not None, shape)


def extract(S):
    out = [np.asarray(S.logEvidence, float), np.asarray(S.marginalizedPosterior, float),
           np.asarray(S.transitionModelDistribution, float), np.asarray(S.localTransitionModelDistribution, float),
           np.concatenate([np.ravel(h) for h in S.hyperParameterDistribution]),
           np.concatenate([np.ravel(e) for e in S.logEvidenceList])]
    if S.storeHistory:
        out += [np.asarray(S.posteriorMeanValues, float), np.asarray(S.posteriorSequence, float),
                np.asarray(S.transitionModelSequence, float)]
    return out


def run(fit):
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
        try:
            return 'ok', fit()
        except Exception as e:  # noqa: BLE001 -- the exception type is what is compared
            return 'exc', type(e).__name__


def close(x, y):
    if x.ndim == 1:  # per-hypothesis evidences: NaN (reference) = dead (product), see the module docstring
        x, y = np.where(np.isnan(x), -np.inf, x), np.where(np.isnan(y), -np.inf, y)
    top = max(1., float(np.nanmax(np.abs(x)))) if x.size and np.isfinite(x).any() else 1.
    return x.shape == y.shape and np.allclose(x, y, rtol=1e-8, atol=1e-13 * top, equal_nan=True)


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
            continue
        R, O = extract(r[1]), extract(o[1])
        if len(R) != len(O) or not all(close(x, y) for x, y in zip(R, O)):
            differ += 1
            print('seed %d DIFF   %s' % (seed, label))
    print('%d cases, %d rejected by the reference itself, %d differ' % (n, rejected, differ))
    return 1 if differ else 0


if __name__ == '__main__':
    sys.exit(main())
