"""Per-CTA timeline of one C2 sweep (BLG_TRACE=<prefix> must be set): python tools/trace_c2.py [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
S = bench.build_study(bl, bench.synthetic_counts(T), 512, bench.GRID, bench.SIGMA_MAX)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
S._executeSweep(sw)
import numpy as np
np.save(os.environ['BLG_TRACE'] + '.radius.npy', sw['program'].host['radius'][:, 0])
