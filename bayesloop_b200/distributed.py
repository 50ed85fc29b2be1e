"""Multi-GPU sharding of the hyper-parameter sweep: one process per GPU, torch.distributed for the plumbing.

The sweep is embarrassingly parallel over hyper-parameter combinations (the reference itself splits this axis over
processes, core.py:1463-1465).  Rows of hyperGridValues are dealt BY PREDICTED COST (deal_by_cost: descending cost,
laps of alternating direction): the cost of a combination grows with its random-walk widths, so a contiguous
np.array_split hands one rank all the wide kernels (measured 82 % scaling efficiency at 2 GPUs) and plain round-robin
still gives every rank a fixed residue of the inner grid index (3 - 5 % at 8 GPUs on a 64 x 64 hyper-grid).  Every rank
runs its own waves with NO per-step communication; merge once:

  * all-gather of the per-combo log-evidence / alive flags (B doubles)                        -- core.py:1336
  * all-reduce(sum) of the per-rank partial local evidence [T]                                 -- core.py:1337,1410
  * all-reduce(max) of the per-rank reference log-weight, local re-base by exp(m_r - M), then
    all-reduce(sum) of the running average [T x G] over NCCL / NVLink                          -- core.py:1339-1340

OnlineStudy deals its hypotheses round-robin (row h of the concatenated hypothesis list -> rank h % world): every
rank filters its own rows of the [H x G] state; per step one all-gather of the H evidence increments (H doubles)
keeps the O(H) bookkeeping of core.py:2171-2215 identical on all ranks, and the marginalised posterior is an
all-reduce(sum) of the per-rank weighted row sums [G] when somebody reads it.

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


def shard_rows(B, rank=None, size=None):
    """Row indices of hyperGridValues owned by a rank: rank, rank + world, rank + 2*world, ..."""
    if rank is None:
        rank, size = world()
    return np.arange(rank, int(B), size)


def deal_by_cost(cost, rank=None, size=None):
    """Row indices owned by a rank when the rows are dealt by predicted cost: rows in descending cost, handed out in
    laps 0 .. size-1, size-1 .. 0, ... (every rank gets one row of each lap, the direction alternates), returned in
    ascending order.  Plain round-robin puts row i on rank i % size, which on a Cartesian hyper-grid whose inner axis
    length is a multiple of the world size gives every rank a fixed residue of the inner index: at C3 (64 x 64 widths
    on 8 GPUs) the last rank owns the widest random walks of every block of 8 (7 % more work than the first; the ranks
    were seen to wait 4.7 % of the fit in the first collective, profiles/r2Y2_bench_n8.json)."""
    if rank is None:
        rank, size = world()
    cost = np.asarray(cost, dtype=float)
    order = np.argsort(-cost, kind='stable')
    pos = np.arange(len(order))
    lap, k = pos // size, pos % size
    owner = np.where(lap % 2 == 0, k, size - 1 - k)
    return np.sort(order[owner == rank])


def gather_rows(eng, logE, alive, Ball, rows=None):
    """All-gather the per-rank log-evidence / alive flags (device tensors) and return them on the host in hyper-grid
    row order.  `rows`: global row index of each of this rank's entries (default: the round-robin deal).  One
    collective of a packed [3 x width] block per rank and one device -> host copy."""
    rank, size = world()
    if rows is None:
        rows = shard_rows(Ball, rank, size)
    rows = np.asarray(rows, dtype=np.int64)
    if size == 1:
        outE, outA = np.empty(int(Ball)), np.zeros(int(Ball), dtype=np.int64)
        outE[rows] = np.asarray(eng.to_host(logE), dtype=float)
        outA[rows] = np.asarray(eng.to_host(alive)).astype(np.int64)
        return outE, outA
    width = -(-int(Ball) // size)
    counts = torch.tensor([len(rows)], dtype=torch.int64, device=logE.device)
    most = counts.clone()
    td.all_reduce(most, op=td.ReduceOp.MAX)
    width = max(width, int(most.item()))
    send = torch.full((3, width), -1.0, dtype=torch.float64, device=logE.device)
    send[0, :logE.shape[0]] = logE
    send[1, :alive.shape[0]] = alive.to(torch.float64)
    send[2, :len(rows)] = torch.from_numpy(rows.astype(np.float64)).to(logE.device)
    recv = [torch.empty_like(send) for _ in range(size)]
    td.all_gather(recv, send)
    host = eng.to_host(torch.stack(recv))
    outE, outA = np.empty(int(Ball)), np.zeros(int(Ball), dtype=np.int64)
    for r in range(size):
        idx = host[r, 2]
        keep = idx >= 0
        outE[idx[keep].astype(np.int64)] = host[r, 0][keep]
        outA[idx[keep].astype(np.int64)] = host[r, 1][keep].astype(np.int64)
    return outE, outA


def gather_dealt(eng, mine, total):
    """All-gather arrays whose FIRST axis is dealt round-robin over the ranks (`shard_rows`) and return the whole
    array in global row order.  `mine`: this rank's rows as a tensor on the engine's device, shape [n_r, ...]."""
    rank, size = world()
    if size == 1:
        return eng.to_host(mine)
    width = -(-int(total) // size)
    tail = tuple(mine.shape[1:])
    send = torch.zeros((width,) + tail, dtype=mine.dtype, device=mine.device)
    send[:mine.shape[0]] = mine
    recv = torch.empty((size, width) + tail, dtype=mine.dtype, device=mine.device)
    td.all_gather(list(recv.unbind(0)), send)  # views of one buffer (gloo's all_gather_into_tensor wants a flat one)
    # ONE device -> host copy per call (an OnlineStudy calls this every step); global row r + size * k sits at [r][k]
    host = np.asarray(eng.to_host(recv))
    return np.swapaxes(host, 0, 1).reshape((width * size,) + tail)[:int(total)].copy()


def min_over_ranks(eng, value):
    """Smallest `value` (a host scalar) of the ranks: for decisions every rank must take alike."""
    if world()[1] == 1:
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=eng.device)
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return float(t.item())


def reduce_sum(eng, tensor):
    if world()[1] > 1:
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def rebase_and_reduce(eng, plan, avg, shift, count):
    """Bring every rank's running average onto the common reference log-weight M = max_r shift_r and sum them
    (the merge of core.py:1339-1340).  `shift` is a device scalar: max all-reduce, device-side re-base
    (blg_rebase), sum all-reduce -- enqueued back to back, no host synchronisation in between."""
    if world()[1] == 1:
        return shift
    top = shift.clone()
    td.all_reduce(top, op=td.ReduceOp.MAX)
    eng.rebase(plan, avg, count, shift, top)
    td.all_reduce(avg, op=td.ReduceOp.SUM)
    return top
