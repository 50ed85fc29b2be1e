"""Experiment: C3-shaped HyperStudy (Gaussian 2-D grid, GaussianRandomWalk on both axes); default dispatch = the
cluster-resident kernels (BLG_NO_CLUSTER2D=1: stream kernels; BLG_TRACE=<prefix>: per-phase cycle counters).
python tools/exp_2d.py [n=256] [T=100] [hyper=12 -> B = hyper^2] [smax=0.1]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesloop_b200 as bl  # noqa: E402
from bayesloop_b200 import engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
smax = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
rng = np.random.default_rng(2)
mu = np.clip(np.cumsum(rng.normal(0, 0.02, T)), -2, 2)
x = rng.normal(mu, 1.0)
eng = E.default_engine()
S = bl.HyperStudy(silent=True)
S.loadData(x, silent=True)
S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n), 'std', bl.oint(0, 3, n)),
      bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, smax, H), target='mean'),
                                    bl.tm.GaussianRandomWalk('s_std', bl.cint(0, smax / 2, H), target='std')),
      silent=True)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
times = {}
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
    t0 = time.perf_counter()
    S._executeSweep(sw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
B = H * H
cells = B * T * n * n
print('grid %dx%d T=%d B=%d: %s  wall %.1f ms  -> %.3g cell-updates/s (fwd+bwd = %d updates)'
      % (n, n, T, B, {k: round(v[-1], 2) for k, v in times.items()}, 1e3 * wall, 2 * cells / wall, 2 * cells), flush=True)
for k, v in times.items():
    per = {'forward': 8, 'backward': 16, 'accumulate': 8}[k]
    print('   %-10s %.2f ms  %.1f GB/s algorithmic  (%.2f us per combo-step at %d CTAs)'
          % (k, v[-1], per * cells / v[-1] / 1e6, 1e3 * v[-1] / T / max(1, -(-B // 148)), min(B, 148)))
