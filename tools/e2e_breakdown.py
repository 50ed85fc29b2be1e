"""Where does the end-to-end time of one HyperStudy.fit() go?  (cProfile on the host side, C2 workload)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import bayesloop_b200 as bl  # noqa: E402

counts = bench.synthetic_counts(10000)


def step():
    S = bench.build_study(bl, counts, 512, 1000, 0.2)
    S.fit(silent=True)
    return S


for _ in range(3):
    S = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    S = step()
torch.cuda.synchronize()
print('e2e ms/fit: %.1f' % (1e3 * (time.perf_counter() - t0) / 3))
pr = cProfile.Profile()
pr.enable()
S = step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
