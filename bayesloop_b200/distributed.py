"""Multi-GPU sharding of the hyper-parameter sweep: one process per GPU, torch.distributed for the plumbing.

The sweep is embarrassingly parallel over hyper-parameter combinations (the reference itself splits this axis over
processes, core.py:1463-1465), so rows of hyperGridValues are split contiguously across ranks exactly like
np.array_split, every rank runs its own waves with NO per-step communication, and the results are merged once:

  * all-gather of the per-combo log-evidence / alive flags (B doubles)                        -- core.py:1336
  * all-reduce(sum) of the per-rank partial local evidence [T]                                 -- core.py:1337,1410
  * all-reduce(max) of the per-rank reference log-weight, local re-base by exp(m_r - M), then
    all-reduce(sum) of the running average [T x G] over NCCL / NVLink                          -- core.py:1339-1340

Backend is NCCL for CUDA engines (fp64 sums over NVSwitch) and gloo for the CPU test harness.
"""
import math

import numpy as np
import torch
import torch.distributed as td


def world():
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_bounds(B):
    """Half-open row range of this rank: the same contiguous split as np.array_split(rows, world_size)[rank]."""
    rank, size = world()
    base, extra = divmod(int(B), size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(eng, logE, alive, Ball):
    """Concatenate the per-rank log-evidence / alive arrays in rank order -> arrays of length Ball on every rank."""
    rank, size = world()
    if size == 1:
        return np.asarray(logE, dtype=float), np.asarray(alive)
    width = -(-int(Ball) // size)
    mine = np.full((2, width), np.nan)
    mine[0, :len(logE)] = logE
    mine[1, :len(alive)] = alive
    send = eng.to_device(mine)
    recv = [torch.empty_like(send) for _ in range(size)]
    td.all_gather(recv, send)
    outE, outA = [], []
    for r, block in enumerate(recv):
        base, extra = divmod(int(Ball), size)
        count = base + (1 if r < extra else 0)
        host = eng.to_host(block)
        outE.append(host[0, :count])
        outA.append(host[1, :count])
    return np.concatenate(outE), np.concatenate(outA).astype(np.int64)


def reduce_sum(eng, tensor):
    if world()[1] > 1:
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def rebase_and_reduce(eng, plan, avg, shift, count):
    """Bring every rank's running average onto the common reference log-weight M = max_r shift_r and sum them."""
    if world()[1] == 1:
        return shift
    top = eng.to_device(np.array([shift if np.isfinite(shift) else -1e308]))
    td.all_reduce(top, op=td.ReduceOp.MAX)
    M = float(eng.to_host(top)[0])
    if np.isfinite(shift) and M > shift:
        eng.scale(plan, avg, count, math.exp(shift - M))
    td.all_reduce(avg, op=td.ReduceOp.SUM)
    return M
