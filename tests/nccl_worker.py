"""Worker of tests/test_gpu_multi.py: one rank per GPU (torchrun), NCCL backend, the CUDA engine.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/nccl_worker.py <out_dir> <case> [<case> ...]

Every rank fits the sharded study through the public API (bayesloop_b200/distributed.py: rows of the hyper-grid /
hypotheses dealt round-robin, merge = all-gather of the evidences + all-reduce of the running average, the
counterpart of the reference's merge in core.py:1335-1340) and writes its results; the test compares EVERY rank with
the golden of the unsharded reference run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    out_dir, names = sys.argv[1], sys.argv[2:]
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    td.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = td.get_rank()
    import bayesloop_b200 as bl
    import parity
    from bayesloop_b200 import engine
    eng = engine.default_engine()
    assert eng.backend.startswith('cuda')
    for name in names:
        before = eng.launch_count()
        S, got = parity.run_case(name, bl)
        assert eng.launch_count() > before
        shard = S._dev['rows'] if type(S).__name__ == 'OnlineStudy' else S.sweepStats['rows']
        np.savez(os.path.join(out_dir, '%s.rank%d.npz' % (name, rank)), shard=np.array(shard), **got)
    td.barrier()
    td.destroy_process_group()


if __name__ == '__main__':
    main()
