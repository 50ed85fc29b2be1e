#!/usr/bin/env python
"""bench.py -- grid-cell updates / second of a HyperStudy sweep (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3                       # this repo's CUDA engine
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus 8 --steps 5 --warmup 3                       # one rank per GPU, NCCL
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # reference CPU path (NumPy port), host cores

Workload (config.workload): BASELINE.json configs[1] = "C2": HyperStudy, Poisson observation model on a 1-D grid
of 1000 rates, GaussianRandomWalk sigma sweep of 512 values PER GPU (weak scaling: 512*N values of the same
interval), synthetic Poisson counts T = 10000 (seed 1).  A "step" is one complete
HyperStudy.fit: forward filter + backward smoother + evidence-weighted averaging of all combinations = 2*B*T*G
grid-cell updates.

  value      sweep with inputs already resident in HBM (CUDA events, max over ranks)
  e2e        the same through the public API bl.HyperStudy(...).fit() with HOST (NumPy) inputs and results:
             host->device copies of data/program and device->host copies of the averaged posterior sequence,
             means and evidences inside the timed region
  roofline   dominant kernel: algorithmic HBM bytes per launch / live CUDA-event duration, vs MEASURED_PEAKS.json
  roofline.fp64 the binding unit of this workload: convolution flop per pass / kernel time vs the measured DFMA peak
  cpu_baseline  oracle/np_oracle.py (NumPy+SciPy port of the reference loop) on a bounded sample, 1 core
  extra.c2_narrow   the same sweep with sigma <= 0.05 (radius <= 17): the HBM-leaning regime (N = 1 only)
  extra.c3_sample   bounded sample of BASELINE.json configs[2] (256 x 256 grid) on the cluster-resident kernels (N = 1)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 1000
T_FULL = 10000
COMBOS_PER_GPU = 512
SIGMA_MAX = 0.2  # "reference-like" sweep of SURVEY.md 8d: sigma_n <= 16.7 grid cells, kernel radius <= 67
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def synthetic_counts(T, seed=1):
    rng = np.random.default_rng(seed)
    lam = 3.0 + 2.0 * np.sin(2.0 * np.pi * np.arange(T) / 2000.0)
    return rng.poisson(lam).astype(np.float64)


def build_study(bl, counts, n_sigma, grid, sigma_max, engine=None):
    S = bl.HyperStudy(silent=True, engine=engine)
    S.loadData(counts, silent=True)
    S.set(bl.om.Poisson('rate', bl.oint(0, 12, grid)),
          bl.tm.GaussianRandomWalk('sigma', bl.cint(0, sigma_max, n_sigma), target='rate'), silent=True)
    return S


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(smax)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this workload, from the committed
    `ncu --set full` capture (profiles/r1_ncu_summary.json); None if the capture does not cover it."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_ncu_summary.json')) as f:
            entry = json.load(f)['bench_c2'][kernel]
        return float(entry['dram_bytes_read']) + float(entry['dram_bytes_write'])
    except Exception:
        return None


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def cpu_port_sample(counts, grid, sigma_max, n_sigma_total, rows, T_cpu):
    """Reference-style NumPy loop (oracle/np_oracle.py) on `rows` of the sweep and the first T_cpu data points.
    Returns (cell updates, seconds).  The lowering comes from the product's host code; the arithmetic timed here
    is NumPy/SciPy only."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import bayesloop_b200 as bl
    import helpers
    import np_oracle
    S = build_study(bl, counts[:T_cpu], n_sigma_total, grid, sigma_max)
    ops, hp, _ = helpers.lowered(S)
    pb = helpers.np_problem(S)
    t0 = time.perf_counter()
    np_oracle.hyper_fit(pb, ops, hp, rows=rows)
    return np_oracle.cell_updates(pb, len(rows)), time.perf_counter() - t0


def _cpu_worker(args):
    counts, grid, sigma_max, n_total, rows, T_cpu = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    return cpu_port_sample(counts, grid, sigma_max, n_total, rows, T_cpu)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (NumPy port; the reference itself is Python and is not installed on
    the GPU box) on all host cores, rows of the sweep split across worker processes like HyperStudy._parallelFit
    (core.py:1463-1465).  Each step is a bounded sample of the C2 workload."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, (os.cpu_count() or 1))
    workers = min(cores, 64)
    n_total = COMBOS_PER_GPU * args.gpus
    T_cpu = min(args.cpu_T, args.T)
    per_worker = 4
    rows_all = np.unique(np.linspace(0, n_total - 1, workers * per_worker).round().astype(int))
    chunks = [list(c) for c in np.array_split(rows_all, workers) if len(c)]
    counts = synthetic_counts(args.T)
    jobs = [(counts, args.grid, args.sigma_max, n_total, rows, T_cpu) for rows in chunks]
    ctx = mp.get_context('spawn')
    times = []
    with ctx.Pool(len(chunks)) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    updates = float(sum(r[0] for r in res))
    ms = 1e3 * float(np.mean(times))
    value = updates / (ms / 1e3)
    sample = '%d of %d sigma values (evenly spaced), first %d of %d time steps, %d worker processes' % (
        len(rows_all), n_total, T_cpu, args.T, len(chunks))
    line = {
        'impl': 'reference', 'metric': 'grid_cell_updates_per_s', 'value': value, 'unit': 'cell-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, n_total),
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': len(chunks), 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def c3_sample(bl, eng, torch, steps=3, warmup=2, n=256, T=200, hyper=6):
    """Secondary measurement (not the headline): a bounded sample of BASELINE.json configs[2] ("C3": Gaussian
    256 x 256 grid, GaussianRandomWalk on both parameters) -- hyper x hyper combos spread evenly over the 64 x 64
    hyper-grid of SURVEY.md 8d, first T time steps, full fit on the cluster-resident 2-D kernels."""
    rng = np.random.default_rng(2)
    mu = np.clip(np.cumsum(rng.normal(0, 0.02, T)), -2, 2)
    sd = np.clip(1.0 + np.cumsum(rng.normal(0, 0.01, T)), 0.5, 2.0)
    x = rng.normal(mu, sd)
    S = bl.HyperStudy(silent=True)
    S.loadData(x, silent=True)
    S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, n), 'std', bl.oint(0, 3, n)),
          bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, hyper), target='mean'),
                                        bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, hyper), target='std')),
          silent=True)
    S._formatData()
    S._createHyperGrid(silent=True)
    sw = S._prepareSweep(False, False)
    ms = {'forward': [], 'backward': [], 'accumulate': []}
    names = {}
    events = []
    plain = eng.run

    def timed(which, plan, flags, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plain(which, plan, flags, **kw)
        e1.record()
        names[which] = eng.last_kernel() if which != 'accumulate' else 'accumulate_kernel'
        events.append((which, e0, e1))

    eng.run = timed
    try:
        for _ in range(warmup):
            S._executeSweep(sw)
        torch.cuda.synchronize()
        del events[:]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            res = S._executeSweep(sw)
        b.record()
        torch.cuda.synchronize()
    finally:
        eng.run = plain
    for which, e0, e1 in events:
        if which in ms:
            ms[which].append(e0.elapsed_time(e1))
    B = hyper * hyper
    cells = float(B) * T * n * n
    step_ms = a.elapsed_time(b) / steps
    per = {'forward': 8.0, 'backward': 16.0, 'accumulate': 8.0}
    kern = {names.get(k, k): {'ms': float(np.mean(v)), 'GBps': per[k] * cells / (float(np.mean(v)) * 1e-3) / 1e9}
            for k, v in ms.items() if v}
    return {'workload': 'C3 sample: Gaussian 2-D grid %dx%d, GRW on both parameters, %dx%d of the 64x64 hyper-grid, '
                        'first %d time steps, full fit' % (n, n, hyper, hyper, T),
            'value': 2.0 * cells / (step_ms * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': step_ms, 'kernels': kern,
            'combos': B, 'log_evidence_best_combo': float(np.max(res[1]))}


def workload_config(args, n_total):
    return {'workload': 'C2 HyperStudy: Poisson 1-D grid=%d, GaussianRandomWalk sigma sweep cint(0,%g,%d) '
                        '(%d per GPU), synthetic counts T=%d, full fit (forward+backward+averaging)'
                        % (args.grid, args.sigma_max, n_total, n_total // max(1, args.gpus), args.T),
            'grid': args.grid, 'T': args.T, 'combos': n_total, 'combos_per_gpu': n_total // max(1, args.gpus),
            'sigma_max': args.sigma_max, 'parallelism': 'combos sharded over %d rank(s)' % args.gpus,
            'l2': 'working set per step (alpha sequences, %.1f GB/GPU) exceeds the 126 MB L2; no explicit flush'
                  % (n_total // max(1, args.gpus) * args.T * args.grid * 8 / 1e9)}


def main():
    global COMBOS_PER_GPU
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--T', type=int, default=T_FULL)
    ap.add_argument('--grid', type=int, default=GRID)
    ap.add_argument('--sigma-max', dest='sigma_max', type=float, default=SIGMA_MAX)
    ap.add_argument('--combos-per-gpu', dest='combos', type=int, default=COMBOS_PER_GPU)
    ap.add_argument('--cpu-T', dest='cpu_T', type=int, default=4000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-c3', dest='no_c3', action='store_true', help='skip the secondary C3 (2-D) sample')
    ap.add_argument('--no-narrow', dest='no_narrow', action='store_true', help='skip the secondary narrow C2 sweep')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    COMBOS_PER_GPU = args.combos

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as td
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        td.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node == --gpus'
    import bayesloop_b200 as bl
    from bayesloop_b200 import engine as eng_mod
    eng = eng_mod.default_engine()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    n_total = args.combos * world
    counts = synthetic_counts(args.T)
    G, T = args.grid, args.T
    updates = 2.0 * n_total * T * G

    # ---- per-kernel event timing (live, on the launching stream) ---------------------------------------------
    kernel_ms = {'forward': [], 'backward': [], 'accumulate': []}
    recording = {'on': False, 'events': []}
    plain_run = eng.run

    kernel_names = {'accumulate': 'accumulate_kernel'}

    def timed_run(which, plan, flags, **kw):
        if recording['on'] and which in kernel_ms:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plain_run(which, plan, flags, **kw)
            e1.record()
            if which != 'accumulate':
                kernel_names[which] = eng.last_kernel() + '_kernel'
            recording['events'].append((which, e0, e1))
        else:
            plain_run(which, plan, flags, **kw)

    eng.run = timed_run

    # ---- device-resident sweep ("value") -----------------------------------------------------------------------
    S = build_study(bl, counts, n_total, G, args.sigma_max)
    S._formatData()
    S._createHyperGrid(silent=True)
    sw = S._prepareSweep(False, False)
    for _ in range(args.warmup):
        S._executeSweep(sw)
    launches0 = eng.launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        recording['on'] = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            S._executeSweep(sw)
        e1.record()
        barrier()
        recording['on'] = False
    dev_ms = e0.elapsed_time(e1) / args.steps
    launches = (eng.launch_count() - launches0) // max(1, args.steps)
    for which, a, b in recording['events']:
        kernel_ms[which].append(a.elapsed_time(b))
    waves = S.sweepStats['waves']
    del sw
    torch.cuda.empty_cache()

    # ---- end to end through the public API ("e2e") --------------------------------------------------------------
    def e2e_step():
        S2 = build_study(bl, counts, n_total, G, args.sigma_max)
        S2.fit(silent=True)
        return S2

    for _ in range(min(args.warmup, 3)):
        S2 = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        S2 = e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    n_local = n_total // world
    h2d = counts.nbytes + n_local * (8 + 4 + 16) + G * 8 + n_local * 8
    d2h = T * G * 8 + T * 8 + n_local * 8 + n_local * 4 + T * 8

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        td.all_reduce(times, op=td.ReduceOp.MAX)
    dev_ms, e2e_ms = [float(x) for x in times.cpu()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        n_loc = n_total // world
        # algorithmic HBM bytes per (combo, time step, cell): forward stores alpha[t] (8 B); backward reads alpha[t] and
        # stores the posterior (16 B); the averaging pass reads every posterior once (8 B)  ->  32 B per cell for the
        # two passes = 16 B per grid-cell update (forward 8 B + backward 24 B in HyperStudy mode)
        bytes_per_cell = {'forward': 8.0, 'backward': 16.0, 'accumulate': 8.0}
        names = {k: kernel_names.get(k, k) for k in kernel_ms}
        per_launch_cells = n_loc * T * G / max(1, waves)
        kern = {}
        for which, samples in kernel_ms.items():
            if samples:
                ms = float(np.mean(samples))
                gbs = bytes_per_cell[which] * per_launch_cells / (ms * 1e-3) / 1e9
                kern[names[which]] = {'ms': ms, 'GBps': gbs, 'frac': gbs / peak, 'bytes_per_cell': bytes_per_cell[which]}
        dominant = max(kern, key=lambda k: kern[k]['ms'])
        total_kernel_ms = sum(v['ms'] for v in kern.values()) * waves
        sweep_gbs = 32.0 * n_loc * T * G / (total_kernel_ms * 1e-3) / 1e9
        # the binding unit of this workload is the FP64 pipe (convolution taps), reported next to the HBM roofline:
        # DFMA per pass = T * G * sum over combos of (2 R_b + 1), R_b = int(4 sigma_b / lattice + 0.5)
        sig = np.asarray(bl.cint(0, args.sigma_max, n_total), dtype=float)[:n_loc]
        lattice = 12.0 / (G + 1)
        radius = np.where(sig > 0, np.floor(4.0 * sig / lattice + 0.5), 0.0)
        taps = np.where(radius > 0, 2.0 * radius + 1.0, 0.0)
        flop_pass = 2.0 * T * G * float(taps.sum())
        fp64_peak = 58.5 * 148 * 1.965e9 * 2 / 1e12  # tools/micro/dfma_bench.cu on B200: 58.5 FMA lanes/clk/SM
        fp64 = {'peak_tflops': fp64_peak, 'peak_source': 'measured: tools/micro/dfma_bench.cu (58.5 DFMA lanes/clk/SM x 148 '
                                                         'SMs x 1.965 GHz)',
                'convolution_flop_per_pass': flop_pass, 'mean_taps': float(taps.mean()),
                'kernels': {name: {'tflops': flop_pass / (v['ms'] * 1e-3) / 1e12,
                                   'frac': flop_pass / (v['ms'] * 1e-3) / 1e12 / fp64_peak}
                            for name, v in kern.items() if 'accumulate' not in name}}
        roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': kern[dominant]['GBps'], 'peak': peak, 'unit': 'GB/s',
                    'frac': kern[dominant]['frac'], 'traffic': ncu_traffic(dominant), 'peak_source': peak_src,
                    'kernels': kern, 'all_kernels_GBps': sweep_gbs, 'all_kernels_frac': sweep_gbs / peak,
                    'kernel_share_of_step': total_kernel_ms / dev_ms, 'fp64': fp64,
                    'note': 'reference-like sigma sweep (kernel radius <= 67): the convolution makes the passes FP64-FMA '
                            'bound, not HBM bound (SURVEY.md 8d); see profiles/ for the FP64 pipe utilisation'}
        line = {
            'metric': 'grid_cell_updates_per_s', 'value': updates / (dev_ms * 1e-3), 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dev_ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, n_total),
            'clocks': clocks.summary(),
            'e2e': {'value': updates / (e2e_ms * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
            'gpu_launches': int(launches) * args.steps,
            'roofline': roofline,
            'log_evidence': float(S2.logEvidence),
        }
        if world == 1 and not args.no_c3 and not args.no_narrow:
            # secondary: the narrow sweep of SURVEY.md 8d (sigma <= 0.05, radius <= 17): the regime where the passes
            # move towards the HBM roofline (fewer taps per stored byte); device-resident, same shape otherwise
            try:
                Sn = build_study(bl, counts, n_total, G, 0.05)
                Sn._formatData()
                Sn._createHyperGrid(silent=True)
                swn = Sn._prepareSweep(False, False)
                kernel_ms_n = {'forward': [], 'backward': [], 'accumulate': []}
                evs = []

                def timed_n(which, plan, flags, **kw):
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    plain_run(which, plan, flags, **kw)
                    a1.record()
                    evs.append((which, a0, a1))

                eng.run = timed_n
                for _ in range(2):
                    Sn._executeSweep(swn)
                del evs[:]
                n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0.record()
                for _ in range(3):
                    Sn._executeSweep(swn)
                n1.record()
                torch.cuda.synchronize()
                eng.run = timed_run
                for which, a0, a1 in evs:
                    if which in kernel_ms_n:
                        kernel_ms_n[which].append(a0.elapsed_time(a1))
                ms_n = n0.elapsed_time(n1) / 3
                cells = float(n_loc) * T * G
                line.setdefault('extra', {})['c2_narrow'] = {
                    'workload': 'C2 with the narrow sweep cint(0,0.05,%d): kernel radius <= 17' % n_total,
                    'value': 2.0 * cells / (ms_n * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': ms_n,
                    'kernels': {k: {'ms': float(np.mean(v)),
                                    'hbm_frac': bytes_per_cell[k] * cells / (float(np.mean(v)) * 1e-3) / 1e9 / peak}
                                for k, v in kernel_ms_n.items() if v}}
                del swn, Sn
                torch.cuda.empty_cache()
            except Exception as exc:
                eng.run = timed_run
                line.setdefault('extra', {})['c2_narrow'] = {'error': repr(exc)}
        if world == 1 and not args.no_c3:
            try:
                line.setdefault('extra', {})['c3_sample'] = c3_sample(bl, eng, torch)
            except Exception as exc:  # secondary measurement: never lose the headline line over it
                line.setdefault('extra', {})['c3_sample'] = {'error': repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            rows = list(np.unique(np.linspace(0, n_total - 1, 24).round().astype(int)))
            T_cpu = min(args.cpu_T, T)
            upd, sec = cpu_port_sample(counts, G, args.sigma_max, n_total, rows, T_cpu)
            line['cpu_baseline'] = {'value': upd / sec, 'unit': 'cell-updates/s', 'cores': 1, 'kind': 'port',
                                    'sample': '%d of %d sigma values (evenly spaced), first %d of %d time steps, '
                                              'oracle/np_oracle.py (NumPy + scipy.ndimage.gaussian_filter1d), %.1f s'
                                              % (len(rows), n_total, T_cpu, T, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


if __name__ == '__main__':
    main()
