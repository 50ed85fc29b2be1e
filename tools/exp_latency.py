"""Experiment: forward/backward kernel time vs number of resident CTAs per SM and kernel radius (not a bench)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402
from bayesloop_b200 import engine as E  # noqa: E402

T = int(os.environ.get('EXP_T', 500))
counts = bench.synthetic_counts(T)
eng = E.default_engine()


def run(B, sigma_lo, sigma_hi, full=False):
    S = bl.HyperStudy(silent=True)
    S.loadData(counts, silent=True)
    S.set(bl.om.Poisson('rate', bl.oint(0, 12, 1000)),
          bl.tm.GaussianRandomWalk('sigma', np.linspace(sigma_lo, sigma_hi, B), target='rate'), silent=True)
    S._formatData()
    S._createHyperGrid(silent=True)
    sw = S._prepareSweep(False, not full)
    times = {'forward': [], 'backward': []}
    plain = eng.run

    def timed(which, plan, flags, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plain(which, plan, flags, **kw)
        e1.record()
        torch.cuda.synchronize()
        times.setdefault(which, []).append(e0.elapsed_time(e1))

    eng.run = timed
    for _ in range(3):
        S._executeSweep(sw)
    eng.run = plain
    f = np.mean(times['forward'][1:])
    b = np.mean(times['backward'][1:]) if times['backward'] else float('nan')
    print('B=%4d sigma=[%.3f,%.3f] full=%d  fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)'
          % (B, sigma_lo, sigma_hi, full, f, 1e3 * f / T, b, 1e3 * b / T), flush=True)


print('env:', {k: v for k, v in os.environ.items() if k.startswith('BLG_')}, flush=True)
if os.environ.get('EXP_MODE', 'all') == 'all':
    for B in (148, 592):
        for s in (0.0, 0.05, 0.2):
            run(B, s, s + 1e-9, full=True)
else:
    run(148, 0.0, 1e-9, full=True)
    run(148, 0.2, 0.2 + 1e-9, full=True)
run(512, 0.0, 0.2, full=True)
run(512, 0.0, 0.05, full=True)
