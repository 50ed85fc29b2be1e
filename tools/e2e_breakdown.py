"""Where does the end-to-end time of one HyperStudy.fit() go?  (C2 workload, wall-clock per host call, per fit)

    python tools/e2e_breakdown.py [fits=5]
Every Engine method is wrapped with a wall-clock timer (no extra synchronisation: blocking calls show the GPU time
they wait for); the previous study is kept alive / dropped exactly as bench.py's e2e loop does."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402
from bayesloop_b200 import engine as E  # noqa: E402

fits = int(sys.argv[1]) if len(sys.argv) > 1 else 5
import argparse  # noqa: E402
wl = bench.C2(argparse.Namespace(T=10000, grid=1000, sigma_max=0.2, combos=512), 1)
acc = collections.OrderedDict()


def wrap(cls, name):
    plain = getattr(cls, name)

    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return plain(*a, **k)
        finally:
            key = name if name != 'run' else 'run:' + a[1]
            acc[key] = acc.get(key, 0.0) + time.perf_counter() - t0

    setattr(cls, name, timed)


for n in ('to_host', 'to_device', 'empty', 'zeros', 'full', 'plan', 'run', 'finalize', 'mix', 'scale', 'free_bytes', 'wave_weights',
          'rebase'):
    wrap(E.Engine, n)
wrap(E.Plan, '__del__')
wrap(bl.HyperStudy, '_prepareSweep')
wrap(bl.HyperStudy, '_executeSweep')
wrap(bl.HyperStudy, '_createHyperGrid')
wrap(bl.HyperStudy, '_formatData')

S = None
for it in range(fits):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    S2 = wl.study(bl)
    t1 = time.perf_counter()
    S2.fit(silent=True)
    t2 = time.perf_counter()
    S = S2  # the previous study (and its pinned result buffer) is released here
    del S2
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print('fit %d: total %.1f ms = build %.1f + fit %.1f + release %.1f + sync %.1f' %
          (it, 1e3 * (t4 - t0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
    print('    ' + '  '.join('%s %.1f' % (k, 1e3 * v) for k, v in acc.items()), flush=True)
