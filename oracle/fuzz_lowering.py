"""TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).  Randomised differential test of the host logic
that is hardest to get right -- the lowering of arbitrary transition-model trees (Combined / Serial with change- and
break-points, nested) into flat operator programs with per-combo windows -- against the unmodified reference:

    python oracle/fuzz_lowering.py [n_cases=150] [seed=0]

Every case draws a study type, an observation model, a random model tree, data with optional gaps and time stamps,
runs both implementations (product: host logic -> C ABI -> CPU oracle) and compares log-evidences, posterior means and
hyper-parameter distributions to 1e-8.  Prints the seeds of disagreeing cases.
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


sys.path.insert(0, os.path.join(ROOT, 'tests'))
from fuzz_cases import draw_case, extract  # noqa: E402,F401


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ref = ref_shim.import_reference()
    from bayesloop_b200 import engine
    engine.set_default_engine(engine.Engine(os.path.join(HERE, 'libblgrid_oracle.so'), 'cpu'))
    import bayesloop_b200 as ours

    def run(bl, build):
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink), np.errstate(all='ignore'):
            try:
                return ('ok', extract(build(bl)))
            except Exception as e:  # noqa: BLE001
                return ('exc', type(e).__name__ + ': ' + str(e)[:100])

    bad = skipped = 0
    for seed in range(seed0, seed0 + n):
        build, label = draw_case(seed)
        r = run(ref, build)
        if r[0] == 'exc':  # configurations the reference itself rejects are not informative
            skipped += 1
            continue
        o = run(ours, build)
        same = o[0] == 'ok' and len(o[1]) == len(r[1]) and all(
            x.shape == y.shape and np.allclose(x, y, rtol=1e-8, atol=1e-12 * max(1.0, float(np.nanmax(np.abs(x[np.isfinite(x)]), initial=0.0))),
                                            equal_nan=True)
            for x, y in zip(r[1], o[1]))
        if not same:
            bad += 1
            print('seed %d DIFF  %s   product: %s' % (seed, label, o[1] if o[0] == 'exc' else 'values differ'))
    print('%d cases, %d rejected by the reference itself, %d differ' % (n, skipped, bad))
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
