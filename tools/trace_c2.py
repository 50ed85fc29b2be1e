"""Per-CTA / per-warp timeline of one C2 sweep (BLG_TRACE=<prefix> must be set): python tools/trace_c2.py [T]
Writes <prefix>.<kernel>.<n>.csv (per CTA: SM, combo, start/end ns; per warp: cycles in the convolution, the epilogue and
at the step barrier, hardware warp id) and <prefix>.radius.npy."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
args = argparse.Namespace(T=T, grid=1000, sigma_max=0.2, combos=512)
wl = bench.C2(args, 1)
S = wl.study(bl)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
S._executeSweep(sw)
np.save(os.environ['BLG_TRACE'] + '.radius.npy', sw['program'].host['radius'][:, 0])
