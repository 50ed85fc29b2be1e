"""Multi-GPU parity on hardware (-m gpu, needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
NCCL-sharded fits of a HyperStudy, a ChangepointStudy and an OnlineStudy against the goldens of the unsharded,
unmodified reference.  What must be reproduced is the reference's own merge of its process-parallel sweep
(core.py:1335-1340: concatenated evidences, logaddexp of the partial averages)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import parity
from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu

CASES = ['syn_hyper_gauss_2d', 'syn_hyper_poisson_sweep', 'syn_cps_gauss_2d', 'syn_online_mixed', 'ref_online_2tm']


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('world', [2])
def test_nccl_ranks_reproduce_the_goldens(world, tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'nccl_worker.py'), str(tmp_path)] + CASES
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    for name in CASES:
        want = load_golden(name)
        shards = []
        for rank in range(world):
            got = dict(np.load(os.path.join(str(tmp_path), '%s.rank%d.npz' % (name, rank))))
            shards.append(list(got.pop('shard')))
            parity.compare(name, got, want, rtol=1e-6, atol_post=1e-12)
            if np.isfinite(float(want['logEvidence'])):
                assert abs(float(got['logEvidence']) - float(want['logEvidence'])) <= 1e-9 * abs(float(want['logEvidence']))
        assert sorted(sum(shards, [])) == list(range(len(sum(shards, [])))) and all(len(s) > 0 for s in shards)
