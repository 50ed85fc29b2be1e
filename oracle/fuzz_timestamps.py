"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Randomised differential test of the time-stamp
logic of Serial models (reference transitionModels.py:289-314, :756-818): IRREGULAR
float time stamps (steps of 0.25 ... 3, offsets, optionally rounded), change- and break-points placed ON a stamp, BETWEEN
two stamps or OUTSIDE the series, random sub-models per segment, optionally wrapped in a Combined model, Study or
HyperStudy (lists of widths / change-point times), segment length 1 (Poisson) or 2 (ScaledAR1: stamps shift by one).
The product lowers all of this to half-open step-index windows (`LoweringContext.steps_equal / steps_between`); the
reference compares time stamps at every step.  Same results to 1e-8, or the same exception type.

    python oracle/fuzz_timestamps.py [n_cases=300] [seed=0]
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
    T = int(rng.integers(5, 13))
    two = rng.random() < 0.3  # segment length 2
    n = T + 1 if two else T
    stamps = np.cumsum(rng.choice([0.25, 0.5, 1.0, 1.5, 3.0], n)) + float(rng.choice([0., -5., 1900.]))
    if rng.random() < 0.3:  # integer-valued stamps, kept strictly increasing
        stamps = np.round(stamps)
        stamps = np.maximum.accumulate(stamps) + np.arange(n) * (0 if len(set(stamps)) == n else 1)
    data = rng.normal(0, 1, n) if two else rng.poisson(3, n).astype(float)
    if two:
        om = lambda bl: bl.om.ScaledAR1('rho', bl.oint(-1, 1, 8), 'sig', bl.oint(0, 3, 7))  # noqa: E731
    else:
        om = lambda bl: bl.om.Poisson('r', bl.oint(0, 8, 12))  # noqa: E731
    target = 'rho' if two else 'r'
    formatted = stamps[1:] if two else stamps  # core.py:350

    def point():
        r = rng.random()
        if r < 0.5:
            return float(rng.choice(formatted))
        if r < 0.8:
            i = int(rng.integers(0, len(formatted) - 1))
            return float(0.5 * (formatted[i] + formatted[i + 1]))
        return float(rng.choice([formatted[0] - 2., formatted[-1] + 2., formatted[0], formatted[-1]]))

    count = [0]

    def leaf(hyper, segment=True):  # a ChangePoint as a SEGMENT of a Serial model would count as one of its points
        kind = str(rng.choice(['static', 'grw', 'grw', 'rs', 'ne', 'indep'] if segment else ['static', 'grw', 'rs', 'ne', 'indep', 'cp']))
        count[0] += 1
        name = 'h%d' % count[0]
        if kind == 'grw':
            if hyper and rng.random() < 0.5:
                value = sorted(set(float(x) for x in np.round(rng.uniform(0, 0.5, 2), 3)))
            else:
                value = float(np.round(rng.uniform(0, 0.5), 3))
            return lambda bl: bl.tm.GaussianRandomWalk(name, value, target=target)
        if kind == 'rs':
            return lambda bl: bl.tm.RegimeSwitch(name, -5.)
        if kind == 'ne':
            return lambda bl: bl.tm.NotEqual(name, -4.)
        if kind == 'indep':
            return lambda bl: bl.tm.Independent()
        if kind == 'cp':
            value = sorted(set([point(), point(), point()])) if hyper and rng.random() < 0.5 else point()
            return lambda bl: bl.tm.ChangePoint(name, value)
        return lambda bl: bl.tm.Static()

    hyper = rng.random() < 0.5
    points = sorted(set(point() for _ in range(int(rng.integers(1, 4)))))
    is_change = [rng.random() < 0.5 for _ in points]
    subs = [leaf(hyper) for _ in range(len(points) + 1)]
    wrap, extra = rng.random() < 0.4, leaf(hyper, segment=False)

    def tree(bl):
        args = [subs[0](bl)]
        for i, (p, change) in enumerate(zip(points, is_change)):
            args.append(bl.tm.ChangePoint('c%d' % i, p) if change else bl.tm.BreakPoint('b%d' % i, p))
            args.append(subs[i + 1](bl))
        serial = bl.tm.SerialTransitionModel(*args)
        return bl.tm.CombinedTransitionModel(extra(bl), serial) if wrap else serial

    kw, r = {}, rng.random()
    if r < 0.15:
        kw = {'forwardOnly': True}
    elif r < 0.3:
        kw = {'evidenceOnly': True}

    def build(bl):
        S = (bl.HyperStudy if hyper else bl.Study)()
        S.loadData(data, timestamps=stamps)
        S.set(om(bl), tree(bl))
        S.fit(**kw)
        return S
    return build, '%s T=%d segment=%d points=%s change=%s wrapped=%s %s' % (
        'HyperStudy' if hyper else 'Study', T, 2 if two else 1, np.round(points, 2), is_change, wrap, kw)


def run(fit):
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
        try:
            return 'ok', fit()
        except Exception as e:  # noqa: BLE001 -- the exception type is what is compared
            return 'exc', type(e).__name__


def extract(S, label):
    out = [np.asarray(S.logEvidence, float)]
    if 'evidenceOnly' not in label and np.isfinite(S.logEvidence):
        out += [np.asarray(S.posteriorMeanValues, float), np.asarray(S.posteriorSequence, float)]
    if getattr(S, 'hyperParameterDistribution', None) is not None and len(getattr(S, 'logEvidenceList', [])) > 0:
        out += [np.asarray(S.logEvidenceList, float), np.asarray(S.hyperParameterDistribution, float)]
    return out


def close(x, y):
    if x.shape != y.shape:
        return False
    if x.ndim >= 2:
        return bool(np.all(np.abs(x - y) <= 1e-8 * np.abs(x) + 1e-12 * np.nanmax(np.abs(x))))
    return np.allclose(x, y, rtol=1e-8, atol=1e-12, equal_nan=True)


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
        R, O = extract(r[1], label), extract(o[1], label)
        if len(R) != len(O) or not all(close(x, y) for x, y in zip(R, O)):
            differ += 1
            print('seed %d DIFF  %r vs %r   %s' % (seed, r[1].logEvidence, o[1].logEvidence, label))
    print('%d cases, %d rejected by the reference itself, %d differ' % (n, rejected, differ))
    return 1 if differ else 0


if __name__ == '__main__':
    sys.exit(main())
