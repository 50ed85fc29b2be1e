"""Experiment: C4-shaped sweep (BASELINE.json configs[3] in miniature): Gaussian 2-D grid n x n, one change-point x
GaussianRandomWalk on both parameters; B = cps * hyper^2 combos.   python tools/exp_c4.py [n=200] [T=300] [cps=10] [hyper=3]
Several GPUs (combos dealt over the ranks, merged over NCCL; strong scaling of the same B):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exp_c4.py 200 300 40 4"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesloop_b200 as bl  # noqa: E402
from bayesloop_b200 import engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
H = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(3)
x = np.concatenate([rng.normal(-0.5, 1.0, T // 2), rng.normal(1.0, 1.0, T - T // 2)])
world = int(os.environ.get('WORLD_SIZE', '1'))
if world > 1:
    import torch.distributed as td
    td.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
eng = E.default_engine()
S = bl.HyperStudy(silent=True)
S.loadData(x, silent=True)
S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n), 'std', bl.oint(0, 3, n)),
      bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', list(np.linspace(T // 10, T - T // 10, cps).astype(int))),
                                    bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, H), target='mean'),
                                    bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, H), target='std')),
      silent=True)
S._formatData()
S._createHyperGrid(silent=True)
sw = S._prepareSweep(False, False)
for _ in range(2):
    S._executeSweep(sw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    res = S._executeSweep(sw)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
if world > 1:
    worst = torch.tensor([dt], device=eng.device, dtype=torch.float64)
    td.all_reduce(worst, op=td.ReduceOp.MAX)
    dt = float(worst.item())
B = cps * H * H
if int(os.environ.get('RANK', '0')) == 0:
    print('grid %dx%d T=%d B=%d (%d change-points x %dx%d sigmas) on %d GPU(s): %.1f ms per full fit -> %.3g cell-updates/s '
          '(kernel: %s), best combo logE %.4f' % (n, n, T, B, cps, H, H, world, 1e3 * dt, 2.0 * B * T * n * n / dt,
                                                  eng.last_kernel(), float(np.max(res[1]))), flush=True)
if world > 1:
    td.destroy_process_group()
