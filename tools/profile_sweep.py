"""Short C2-shaped sweep for ncu captures (never a bench number): python tools/profile_sweep.py [T] [combos] [sigma_max]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
smax = float(sys.argv[3]) if len(sys.argv) > 3 else bench.SIGMA_MAX
S = bench.build_study(bl, bench.synthetic_counts(T), B, bench.GRID, smax)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
for _ in range(2):
    S._executeSweep(sw)
print('logE of the first / last combo:', S.sweepStats, flush=True)
