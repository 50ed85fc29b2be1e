"""Experiment: C5-shaped OnlineStudy (ScaledAR1 2-D grid n x n, 240 GRW-pair hypotheses + 15 RegimeSwitch + 1
Independent = 256) -- seconds per step() through the public API.   python tools/exp_online.py [n=512] [steps=40]
Several GPUs (hypotheses dealt over the ranks, one all-gather of 256 doubles per step):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exp_online.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesloop_b200 as bl  # noqa: E402
from bayesloop_b200 import engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(4)
x = np.zeros(steps + 12)
for i in range(1, len(x)):
    x[i] = 0.6 * x[i - 1] + rng.normal(0, 1.0)
world = int(os.environ.get('WORLD_SIZE', '1'))
if world > 1:
    import torch.distributed as td
    td.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
eng = E.default_engine()
S = bl.OnlineStudy(storeHistory=False, silent=True)
S.setOM(bl.om.ScaledAR1('rho', bl.oint(-1, 1, n), 'sigma', bl.oint(0, 3, n)), silent=True)
S.add('normal', bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s1', bl.cint(0, 0.03, 16), target='rho'),
                                              bl.tm.GaussianRandomWalk('s2', bl.cint(0, 0.03, 15), target='sigma')))
S.add('chaotic', bl.tm.RegimeSwitch('p', bl.cint(-10, -3, 15)))
S.add('indep', bl.tm.Independent())
for d in x[:12]:
    S.step(d)
torch.cuda.synchronize()
t0 = time.perf_counter()
for d in x[12:]:
    S.step(d)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
if world > 1:
    worst = torch.tensor([dt], device=eng.device, dtype=torch.float64)
    td.all_reduce(worst, op=td.ReduceOp.MAX)
    dt = float(worst.item())
if int(os.environ.get('RANK', '0')) == 0:
    print('grid %dx%d, 256 hypotheses on %d GPU(s): %.2f ms per step -> %.3g cell-updates/s (kernel: %s)'
          % (n, n, world, 1e3 * dt, 256.0 * n * n / dt, eng.last_kernel()), flush=True)
post = S.marginalizedPosterior  # collective read with several ranks
if int(os.environ.get('RANK', '0')) == 0:
    # fingerprints for comparing kernel families (BLG_ONLINE2D=1 vs the stream kernels) on the same stream
    print('logE %.12f  mean rho %.12f  mean sigma %.12f  P(normal) %.12f'
          % (S.logEvidence, S.getCurrentParameterMeanValue('rho'), S.getCurrentParameterMeanValue('sigma'),
             S.getCurrentTransitionModelProbability('normal')), flush=True)
if world > 1:
    td.destroy_process_group()
