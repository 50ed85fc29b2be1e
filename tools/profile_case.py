"""ncu target: one HyperStudy sweep with B combos at a single sigma (not a bench): profile_case.py B sigma T"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402

B, sigma, T = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
S = bl.HyperStudy(silent=True)
S.loadData(bench.synthetic_counts(T), silent=True)
S.set(bl.om.Poisson('rate', bl.oint(0, 12, 1000)),
      bl.tm.GaussianRandomWalk('sigma', np.linspace(sigma, sigma + 1e-9, B), target='rate'), silent=True)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
for _ in range(2):
    S._executeSweep(sw)
print('done', S.sweepStats, flush=True)
