"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Randomised differential test of the change-point
prefix sharing: random (change-points) x (hyper-parameter lists) sweeps -- Poisson / Gaussian /
ScaledAR1 grids, one or two random walks, optional RegimeSwitch / NotEqual / Static in any order, hyper-priors of every
kind (none / callable / array / SymPy random variable), missing values, time
stamps with an offset, forwardOnly / evidenceOnly, random wave caps -- fitted by the unmodified reference, by the product
with the shared schedule and by the product with the plain schedule (host logic -> C ABI -> CPU oracle); log-evidences,
hyper-parameter distributions, posterior means, averaged posterior sequences and local evidences must agree to 1e-8.

    python oracle/fuzz_sharing.py [n_cases=300] [seed=0]

The local evidence is left out when a combination died: the reference then holds np.empty leftovers (core.py:360, :400).
Found so far: shared passes next to range-limited operators (Serial segments) and the local evidences of dead
combinations -- both fixed in bayesloop_b200/core.py (_share_structure, _executeSharedSweep).
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

ref = ref_shim.import_reference()
from bayesloop_b200 import engine  # noqa: E402

engine.set_default_engine(engine.Engine(os.path.join(HERE, 'libblgrid_oracle.so'), 'cpu'))
import bayesloop_b200 as ours  # noqa: E402

def draw(seed):
    """One random sweep: returns build(bl, share=None) and a label."""
    rng = np.random.default_rng(seed)
    T = int(rng.integers(6, 16))
    t0 = int(rng.choice([0, 0, 1900]))
    which = str(rng.choice(['poisson', 'gauss', 'ar1']))
    g = rng.integers(6, 14, 2)
    if which == 'poisson':
        data, params = rng.poisson(3, T).astype(float), ['r']
        om = lambda bl: bl.om.Poisson('r', bl.oint(0, 8, int(g[0])))  # noqa: E731
    elif which == 'gauss':
        data, params = rng.normal(0.2, 0.9, T), ['m', 's']
        om = lambda bl: bl.om.Gaussian('m', bl.cint(-3, 3, int(g[0])), 's', bl.oint(0, 3, int(g[1])))  # noqa: E731
    else:
        data, params = rng.normal(0, 1, T + 1), ['rho', 'sig']
        om = lambda bl: bl.om.ScaledAR1('rho', bl.oint(-1, 1, int(g[0])), 'sig', bl.oint(0, 3, int(g[1])))  # noqa: E731
    if rng.random() < 0.3:
        data[int(rng.integers(1, len(data) - 1))] = np.nan
    stamps = None if t0 == 0 else np.arange(t0, t0 + len(data))
    first = t0 + (1 if which == 'ar1' else 0)
    cls = str(rng.choice(['HyperStudy', 'ChangepointStudy']))
    if rng.random() < 0.3:
        cps = 'all'
    else:
        cps = sorted(set(int(x) for x in rng.integers(first - 1, first + T + 1, int(rng.integers(2, 6)))))
    parts = [('cp', cps)]
    for _ in range(int(rng.integers(1, 3))):
        widths = sorted(set(float(x) for x in np.round(rng.uniform(0, 0.6, int(rng.integers(1, 4))), 3)))
        parts.append(('grw', str(rng.choice(params)), widths))
    if rng.random() < 0.4:
        parts.append(('rs', [float(x) for x in rng.choice([-8., -6., -4., -2.], int(rng.integers(1, 3)), replace=False)]))
    if rng.random() < 0.2:
        parts.append(('ne', -3.))
    if rng.random() < 0.2:
        parts.append(('static',))
    parts = [parts[i] for i in rng.permutation(len(parts))]
    # hyper-prior of every part that has hyper-parameters: none / callable / array / SymPy random variable
    # (core.py:1183-1240; the array is normalised in place there, so each build gets a fresh one)
    prior_kinds = [str(rng.choice(['none', 'none', 'callable', 'array', 'sympy'])) for _ in parts]
    prior_seeds = [int(rng.integers(0, 2 ** 31)) for _ in parts]

    def hyper_prior(k, values):
        import sympy.stats
        kind = prior_kinds[k]
        if kind == 'none' or isinstance(values, str):
            return {}
        if kind == 'callable':
            return {'prior': lambda x: 1. / (1. + np.abs(x))}
        if kind == 'array':
            return {'prior': 0.25 + np.random.default_rng(prior_seeds[k]).random(len(values))}
        return {'prior': sympy.stats.Normal('h%d' % k, float(np.mean(values)), 1. + float(np.ptp(values)))}
    wave = None if rng.random() < 0.5 else int(rng.integers(1, 9))
    kw, r = {}, rng.random()
    if r < 0.15:
        kw = {'forwardOnly': True}
    elif r < 0.3:
        kw = {'evidenceOnly': True}

    def tree(bl):
        models = []
        for k, p in enumerate(parts):
            if p[0] == 'cp':
                models.append(bl.tm.ChangePoint('t', p[1], **hyper_prior(k, p[1])))
            elif p[0] == 'grw':
                models.append(bl.tm.GaussianRandomWalk('s%d' % k, p[2], target=p[1], **hyper_prior(k, p[2])))
            elif p[0] == 'rs':
                models.append(bl.tm.RegimeSwitch('p%d' % k, p[1], **hyper_prior(k, p[1])))
            elif p[0] == 'ne':
                models.append(bl.tm.NotEqual('q%d' % k, p[1]))
            else:
                models.append(bl.tm.Static())
        return bl.tm.CombinedTransitionModel(*models) if len(models) > 1 else models[0]

    def build(bl, share=None):
        S = getattr(bl, cls)()
        if stamps is None:
            S.loadData(data)
        else:
            S.loadData(data, timestamps=stamps)
        S.set(om(bl), tree(bl))
        if share is not None:  # product only
            S.shareChangepoints, S.maxWave = share, wave
        S.fit(**kw)
        return S
    return build, '%s/%s T=%d %s priors %s wave=%s %s' % (cls, which, T, [p[0] for p in parts], prior_kinds, wave, kw)


def extract(S, evidence_only):
    out = [np.asarray(S.logEvidence, float), np.asarray(S.logEvidenceList, float), np.asarray(S.hyperParameterDistribution, float)]
    if not evidence_only and np.isfinite(S.logEvidence):
        out += [np.asarray(S.posteriorMeanValues, float), np.asarray(S.posteriorSequence, float)]
        if np.all(np.isfinite(np.asarray(S.logEvidenceList, float))):  # dead combinations: np.empty leftovers in the reference
            out.append(np.asarray(S.localEvidence, float))
    return out


def run(fit):
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
        try:
            return 'ok', fit()
        except Exception as e:  # noqa: BLE001 -- a rejection must be a rejection on both sides
            return 'exc', type(e).__name__ + ': ' + str(e)[:80]


def close(x, y):
    if x.shape != y.shape:
        return False
    if np.allclose(x, y, rtol=1e-8, atol=1e-12 if x.ndim == 0 else 1e-300, equal_nan=True):
        return True
    # posterior sequences: relative to the largest cell
    return x.ndim >= 2 and bool(np.all(np.abs(x - y) <= 1e-8 * np.abs(x) + 1e-13 * np.nanmax(np.abs(x))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    differ = shared = rejected = 0
    for seed in range(seed0, seed0 + n):
        build, label = draw(seed)
        r, a, b = run(lambda: build(ref)), run(lambda: build(ours, True)), run(lambda: build(ours, False))
        if r[0] == 'exc':
            rejected += 1
            if a[0] != 'exc' or b[0] != 'exc':
                differ += 1
                print('seed %d: the reference rejects (%s), the product does not   %s' % (seed, r[1], label))
            continue
        if a[0] == 'exc' or b[0] == 'exc':
            differ += 1
            print('seed %d: product raises %s   %s' % (seed, a[1] if a[0] == 'exc' else b[1], label))
            continue
        shared += bool(a[1].sweepStats.get('shared'))
        evidence_only = 'evidenceOnly' in label
        R, A, B = extract(r[1], evidence_only), extract(a[1], evidence_only), extract(b[1], evidence_only)
        bad = [name for x, y, z in zip(R, A, B) for u, name in ((y, 'shared'), (z, 'plain')) if not close(x, u)]
        if len(R) != len(A) or len(R) != len(B) or bad:
            differ += 1
            print('seed %d DIFF (%s schedule)   %s' % (seed, ', '.join(sorted(set(bad))) or 'result sets', label))
    print('%d cases, %d rejected by the reference itself, %d ran with the shared schedule, %d differ' % (n, rejected, shared, differ))
    return 1 if differ else 0


if __name__ == '__main__':
    sys.exit(main())
