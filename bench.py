#!/usr/bin/env python
"""bench.py -- grid-cell updates / second of the bayesloop sweeps (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus 1 --steps 5 --warmup 3                       # this repo's CUDA engine, headline config C2
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus 8 --steps 5 --warmup 3                       # one rank per GPU, NCCL
    python bench.py --config c3 --gpus 1                                # make another BASELINE config the headline
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # reference CPU path (NumPy port), host cores

Workloads (BASELINE.json configs, inputs of SURVEY.md section 8d; one "step" = one complete fit of the study):

  c2  (default headline, configs[1])  HyperStudy, Poisson 1-D grid 1000, GaussianRandomWalk sigma sweep, T = 10000.
      512 sigma values PER GPU (weak scaling: cint(0, 0.2, 512 N)); full fit = 2 B T G cell updates.
  c3  (configs[2])  HyperStudy, Gaussian 256 x 256 grid, GRW on both parameters, the FULL 64 x 64 hyper-grid (4096
      combinations), steady-state window of the first 200 time steps, STRONG scaling: 4096 / N
      combinations per rank, evidence all-gather + all-reduce of the [T x G] average inside the timed region.
  c4  (configs[3])  change-point study, Gaussian 200 x 200 grid, 100 change-points x 10 x 10 random-walk widths =
      10000 combinations, window of 400 time steps with the change-points spread over it, STRONG scaling.
  c5  (configs[4])  OnlineStudy, ScaledAR1 512 x 512 grid, 256 hypotheses (240 GRW pairs + 15 RegimeSwitch + 1
      Independent), window of 200 steps, hypotheses dealt over the ranks (STRONG), one all-gather per step.

The headline line carries `value` (inputs resident in HBM, CUDA events, max over ranks), `e2e` (public API with HOST
inputs/results), `roofline` (dominant kernel, algorithmic bytes / live event time vs MEASURED_PEAKS.json),
`cpu_baseline` (NumPy port of the reference loop on a bounded sample, 1 core).  With the default headline (c2) the
other three configs are measured too and reported under `extra` at EVERY N, so the driver's 1/2/4/8-GPU runs also
record the strong scaling of c3 / c4 / c5 (each with its collective share).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
FP64_LANES = 63.8          # tools/micro/dmma_mix_bench.cu on B200: FP64 MACs / clk / SM of mma.m8n8k4.f64 (DFMA: 56.9)
BYTES_PER_CELL = {'forward': 8.0, 'backward': 16.0, 'accumulate': 8.0}  # algorithmic HBM bytes


# ------------------------------------------------------------------------------------------------ workloads
def synthetic_counts(T, seed=1):
    rng = np.random.default_rng(seed)
    lam = 3.0 + 2.0 * np.sin(2.0 * np.pi * np.arange(T) / 2000.0)
    return rng.poisson(lam).astype(np.float64)


def gauss_series(T, seed, jump_at=None):
    rng = np.random.default_rng(seed)
    mu = np.clip(np.cumsum(rng.normal(0, 0.02, T)), -2, 2)
    if jump_at is not None:
        mu[jump_at:] += 1.5
    sd = np.clip(1.0 + np.cumsum(rng.normal(0, 0.01, T)), 0.5, 2.0)
    return rng.normal(mu, sd)


def ar1_series(T, seed=4):
    rng = np.random.default_rng(seed)
    x = np.zeros(T)
    for i in range(1, T):
        x[i] = 0.6 * x[i - 1] + rng.normal(0, 1.0)
    return x


class C2:
    key, scaling, kind = 'c2', 'weak', 'sweep'

    def __init__(self, args, world):
        self.T, self.G, self.sigma_max = args.T or 10000, args.grid, args.sigma_max
        self.B = args.combos * world
        self.world = world
        self.data = synthetic_counts(self.T)
        self.shape = (self.G,)

    def study(self, bl, engine=None, T=None, sigma_max=None):
        S = bl.HyperStudy(silent=True, engine=engine)
        S.loadData(self.data[:T or self.T], silent=True)
        S.set(bl.om.Poisson('rate', bl.oint(0, 12, self.G)),
              bl.tm.GaussianRandomWalk('sigma', bl.cint(0, sigma_max or self.sigma_max, self.B), target='rate'),
              silent=True)
        return S

    def describe(self):
        per = self.B // self.world
        return {'workload': 'C2 HyperStudy: Poisson 1-D grid=%d, GaussianRandomWalk sigma sweep cint(0,%g,%d) '
                            '(%d per GPU), synthetic counts T=%d, full fit (forward+backward+averaging)'
                            % (self.G, self.sigma_max, self.B, per, self.T),
                'grid': self.G, 'T': self.T, 'combos': self.B, 'combos_per_gpu': per, 'sigma_max': self.sigma_max,
                'parallelism': 'combos sharded over %d rank(s)' % self.world,
                'l2': 'working set per step (alpha sequences, %.1f GB/GPU) exceeds the 126 MB L2; no explicit flush'
                      % (per * self.T * self.G * 8 / 1e9)}


class C3:
    key, scaling, kind = 'c3', 'strong', 'sweep'

    def __init__(self, args, world):
        self.T, self.n, self.h = args.T or 200, 256, args.hyper
        self.G, self.B, self.world = self.n * self.n, self.h * self.h, world
        self.data = gauss_series(self.T, 2)
        self.shape = (self.n, self.n)

    def study(self, bl, engine=None, T=None):
        S = bl.HyperStudy(silent=True, engine=engine)
        S.loadData(self.data[:T or self.T], silent=True)
        S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, self.n), 'std', bl.oint(0, 3, self.n)),
              bl.tm.CombinedTransitionModel(bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, self.h), target='mean'),
                                            bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, self.h), target='std')),
              silent=True)
        return S

    def describe(self):
        return {'workload': 'C3 HyperStudy: Gaussian 2-D grid %dx%d, GaussianRandomWalk on both parameters, full %dx%d '
                            'hyper-grid (%d combos), steady-state window of %d time steps, full fit'
                            % (self.n, self.n, self.h, self.h, self.B, self.T),
                'grid': [self.n, self.n], 'T': self.T, 'combos': self.B, 'combos_per_gpu': self.B // self.world,
                'parallelism': 'combos dealt over %d rank(s); all-gather of evidences + all-reduce of the [T x G] average'
                               % self.world,
                'l2': 'alpha sequences of a wave (%.1f GB per rank) exceed the 126 MB L2; no explicit flush'
                      % (min(self.B // self.world, 1400) * self.T * self.G * 8 / 1e9)}


class C4(C3):
    key = 'c4'

    def __init__(self, args, world):
        self.T, self.n = args.T or 400, 200
        self.ncp, self.hs = args.changepoints, 10
        self.G, self.B, self.world = self.n * self.n, self.ncp * self.hs * self.hs, world
        self.data = gauss_series(self.T, 3, jump_at=self.T // 2)
        self.shape = (self.n, self.n)

    def study(self, bl, engine=None, T=None):
        T = T or self.T
        S = bl.ChangepointStudy(silent=True, engine=engine)
        S.loadData(self.data[:T], silent=True)
        ncp = min(self.ncp, T - 2)
        step = max(1, T // ncp)
        points = np.arange(step // 2, T - 1, step)[:ncp]
        S.set(bl.om.Gaussian('mean', bl.cint(-3, 3, self.n), 'std', bl.oint(0, 3, self.n)),
              bl.tm.CombinedTransitionModel(bl.tm.ChangePoint('tChange', points),
                                            bl.tm.GaussianRandomWalk('s_mean', bl.cint(0, 0.1, self.hs), target='mean'),
                                            bl.tm.GaussianRandomWalk('s_std', bl.cint(0, 0.05, self.hs), target='std')),
              silent=True)
        return S

    def describe(self):
        return {'workload': 'C4 ChangepointStudy: Gaussian 2-D grid %dx%d, %d change-points x %dx%d random-walk widths '
                            '(%d combos), window of %d time steps with the change-points spread over it, full fit'
                            % (self.n, self.n, self.ncp, self.hs, self.hs, self.B, self.T),
                'grid': [self.n, self.n], 'T': self.T, 'combos': self.B, 'combos_per_gpu': self.B // self.world,
                'parallelism': 'combos dealt over %d rank(s); all-gather of evidences + all-reduce of the [T x G] average'
                               % self.world,
                'l2': 'alpha sequences of a wave exceed the 126 MB L2; no explicit flush'}


class C5:
    key, scaling, kind = 'c5', 'strong', 'online'

    def __init__(self, args, world):
        self.T, self.n = args.T or 200, 512
        self.G, self.B, self.world = self.n * self.n, 256, world
        self.lead = 12
        self.data = ar1_series(self.T + self.lead + 1)
        self.shape = (self.n, self.n)

    def study(self, bl, engine=None, T=None):
        S = bl.OnlineStudy(storeHistory=False, silent=True, engine=engine)
        S.setOM(bl.om.ScaledAR1('rho', bl.oint(-1, 1, self.n), 'sigma', bl.oint(0, 3, self.n)), silent=True)
        with contextlib.redirect_stdout(io.StringIO()):
            S.add('normal', bl.tm.CombinedTransitionModel(
                bl.tm.GaussianRandomWalk('s1', bl.cint(0, 0.03, 16), target='rho'),
                bl.tm.GaussianRandomWalk('s2', bl.cint(0, 0.03, 15), target='sigma')))
            S.add('chaotic', bl.tm.RegimeSwitch('p', bl.cint(-10, -3, 15)))
            S.add('indep', bl.tm.Independent())
        return S

    def describe(self):
        return {'workload': 'C5 OnlineStudy: ScaledAR1 2-D grid %dx%d, 256 hypotheses (240 GRW pairs + 15 RegimeSwitch + '
                            '1 Independent), window of %d step() calls after %d lead-in steps'
                            % (self.n, self.n, self.T, self.lead),
                'grid': [self.n, self.n], 'T': self.T, 'combos': self.B, 'combos_per_gpu': self.B // self.world,
                'parallelism': 'hypotheses dealt over %d rank(s); one all-gather of 256 evidence increments per step'
                               % self.world,
                'l2': 'hypothesis states (%.0f MB per rank) exceed the 126 MB L2; no explicit flush'
                      % (self.B // self.world * self.G * 8 / 1e6)}


WORKLOADS = {'c2': C2, 'c3': C3, 'c4': C4, 'c5': C5}


# ------------------------------------------------------------------------------------------------ helpers
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


def ncu_traffic(config, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this workload from the committed
    `ncu --set full` captures (profiles/*_ncu_summary.json, written by tools/ncu_summary.py); the newest capture that
    covers the kernel wins and its tag is reported, so a stale number is visible as such."""
    best = (None, None)
    pdir = os.path.join(ROOT, 'profiles')
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if not name.endswith('_ncu_summary.json'):
            continue
        try:
            with open(os.path.join(pdir, name)) as f:
                entry = json.load(f).get('bench_' + config, {}).get(kernel)
            if entry:
                best = (float(entry['dram_bytes_read']) + float(entry['dram_bytes_write']), name)
        except Exception:
            continue
    return best


def hbm_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def taps_of(sw):
    """Convolution taps per cell and step of every combo of a prepared sweep (sum over its GRW operators)."""
    radius = np.asarray(sw['program'].host['radius'], dtype=float)
    param = np.asarray(sw['program'].host['param'], dtype=float)
    grw = np.array([k == 1 for k in sw['program'].kinds], dtype=bool)
    if radius.size == 0 or not grw.any():
        return np.zeros(len(radius))
    active = (radius[:, grw] > 0) & (param[:, grw] > 0)
    return np.where(active, 2.0 * radius[:, grw] + 1.0, 0.0).sum(axis=1)


class Timers:
    """CUDA-event brackets around engine calls and collectives, on the launching (current) stream."""

    def __init__(self, torch):
        self.torch, self.on, self.events = torch, False, []

    def wrap(self, label, fn, name_of=None):
        def timed(*a, **kw):
            if not self.on:
                return fn(*a, **kw)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            self.events.append((label(*a, **kw) if callable(label) else label, e0, e1, name_of() if name_of else None))
            return out
        return timed

    def collect(self):
        ms, names = {}, {}
        for label, e0, e1, name in self.events:
            ms.setdefault(label, []).append(e0.elapsed_time(e1))
            if name:
                names[label] = name
        self.events = []
        return ms, names


def measure_sweep(wl, bl, eng, torch, td, world, steps, warmup, local_rank, e2e=True, clocks=False):
    """Device-resident sweep (`value`), per-kernel and per-collective event times, optional end-to-end fit."""
    from bayesloop_b200 import distributed as dist

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    timers = Timers(torch)
    plain_run, plain_apply = eng.run, eng.share_apply
    eng.run = timers.wrap(lambda which, *a, **kw: which, plain_run, name_of=eng.last_kernel)
    eng.share_apply = timers.wrap('share_apply', plain_apply)
    plain = {k: getattr(dist, k) for k in ('rebase_and_reduce', 'gather_rows', 'reduce_sum')}
    for k, fn in plain.items():
        setattr(dist, k, timers.wrap('collective:' + k, fn))
    try:
        S = wl.study(bl)
        S._formatData()
        if hasattr(S, '_prepareChangepoints'):
            S._prepareChangepoints(silent=True)
        else:
            S._createHyperGrid(silent=True)
        sw = S._prepareSweep(False, False)
        taps = taps_of(sw)
        for _ in range(warmup):
            S._executeSweep(sw)
        launches0 = eng.launch_count()
        sampler = ClockSampler(local_rank) if clocks else contextlib.nullcontext()
        with sampler as clk:
            barrier()
            timers.on = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                res = S._executeSweep(sw)
            e1.record()
            barrier()
            timers.on = False
        dev_ms = e0.elapsed_time(e1) / steps
        launches = (eng.launch_count() - launches0) // max(1, steps)
        ms, names = timers.collect()
        stats = dict(S.sweepStats)
        logE_best = float(np.max(res[1])) if len(res[1]) else float('nan')
        n_local = sw['B']
        del sw, res
        torch.cuda.empty_cache()
        e2e_ms, logE = None, None
        if e2e:
            for _ in range(min(warmup, 2)):
                S2 = wl.study(bl)
                S2.fit(silent=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                S2 = wl.study(bl)
                S2.fit(silent=True)
            barrier()
            e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
            logE = float(S2.logEvidence)
            del S2
            torch.cuda.empty_cache()
    finally:
        eng.run, eng.share_apply = plain_run, plain_apply
        for k, fn in plain.items():
            setattr(dist, k, fn)
    times = torch.tensor([dev_ms, e2e_ms or 0.0], dtype=torch.float64, device='cuda')
    if world > 1:
        td.all_reduce(times, op=td.ReduceOp.MAX)
    dev_ms, e2e_max = [float(x) for x in times.cpu()]
    return dict(dev_ms=dev_ms, e2e_ms=e2e_max if e2e else None, ms=ms, names=names, launches=int(launches), stats=stats,
                steps=steps,
                n_local=n_local, taps=taps, logE=logE, logE_best=logE_best,
                clocks=clk.summary() if clocks else None)


def sweep_report(wl, m, peak, peak_src, world):
    """Turn the raw timings of measure_sweep into the bench line's value / roofline / e2e objects."""
    T, G = wl.T, wl.G
    updates = 2.0 * wl.B * T * G
    waves = max(1, m['stats'].get('waves', 1))
    n_loc = m['n_local']
    # change-point prefix sharing skips work: the passes only execute part of the nominal updates (this rank's numbers)
    shared = bool(m['stats'].get('shared'))
    executed = float(m['stats'].get('executed_updates', 2.0 * n_loc * T * G))
    nominal = float(m['stats'].get('nominal_updates', 2.0 * n_loc * T * G))
    done = {'forward': executed / 2.0, 'backward': executed / 2.0, 'accumulate': float(n_loc) * T * G}  # cells per pass
    kern = {}
    for which in ('forward', 'backward', 'accumulate'):
        if which in m['ms']:
            total = float(np.sum(m['ms'][which])) / m['steps']  # ms per sweep (all launches)
            name = (m['names'].get(which, which) + '_kernel') if which != 'accumulate' else 'accumulate_kernel'
            gbs = BYTES_PER_CELL[which] * done[which] / (total * 1e-3) / 1e9
            kern[name] = {'ms': total, 'launches_per_step': len(m['ms'][which]) // m['steps'], 'GBps': gbs, 'frac': gbs / peak,
                          'bytes_per_cell': BYTES_PER_CELL[which]}
    coll = {k.split(':', 1)[1]: float(np.sum(v)) / m['steps'] for k, v in m['ms'].items() if k.startswith('collective:')}
    if 'share_apply' in m['ms']:  # prefix sharing: own rows x shared backward message, 16 B per suffix cell
        total = float(np.sum(m['ms']['share_apply'])) / m['steps']
        kern['share_apply_kernel'] = {'ms': total, 'launches_per_step': len(m['ms']['share_apply']) // m['steps'],
                                      'GBps': 16.0 * executed / 4.0 / (total * 1e-3) / 1e9,
                                      'frac': 16.0 * executed / 4.0 / (total * 1e-3) / 1e9 / peak, 'bytes_per_cell': 16.0}
    dominant = max(kern, key=lambda k: kern[k]['ms'])
    kernel_ms = sum(v['ms'] for v in kern.values())
    coll_ms = sum(coll.values())
    fp64_peak = FP64_LANES * 148 * 1.965e9 * 2 / 1e12
    flop_pass = 2.0 * T * G * float(m['taps'].sum())
    fp64 = {'peak_tflops': fp64_peak,
            'peak_source': 'measured: tools/micro/dmma_mix_bench.cu (63.8 FP64 MAC/clk/SM with mma.m8n8k4.f64 x 148 SMs x 1.965 GHz; '
                           'profiles/r2s_dmma_mix.txt)',
            'convolution_flop_per_pass': flop_pass, 'mean_taps': float(m['taps'].mean()) if len(m['taps']) else 0.0,
            'kernels': {name: {'tflops': flop_pass / (v['ms'] * 1e-3) / 1e12,
                               'frac': flop_pass / (v['ms'] * 1e-3) / 1e12 / fp64_peak}
                        for name, v in kern.items() if 'accumulate' not in name}}
    traffic, tag = ncu_traffic(wl.key, dominant)
    roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': kern[dominant]['GBps'], 'peak': peak, 'unit': 'GB/s',
                'frac': kern[dominant]['frac'], 'traffic': traffic,
                'traffic_source': ('committed ncu capture profiles/%s' % tag) if tag else None,
                'peak_source': peak_src, 'kernels': kern,
                'all_kernels_GBps': 32.0 * n_loc * T * G / (kernel_ms * 1e-3) / 1e9,
                'all_kernels_frac': 32.0 * n_loc * T * G / (kernel_ms * 1e-3) / 1e9 / peak,
                'kernel_share_of_step': kernel_ms / m['dev_ms'], 'collectives_ms': coll,
                'collective_share_of_step': coll_ms / m['dev_ms'], 'waves': waves, 'fp64': fp64}
    out = {'value': updates / (m['dev_ms'] * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': m['dev_ms'],
           'gpu_launches_per_step': m['launches'], 'roofline': roofline}
    if shared:
        # a shortcut that skips work reports nominal and executed updates separately.  `value` counts the
        # NOMINAL updates (what the reference executes for this study); the rooflines above count the executed ones.
        out['sharing'] = {'what': 'change-point prefix sharing (history erased at tChange, transitionModels.py:300-312)',
                          'nominal_updates_per_rank': nominal, 'executed_updates_per_rank': executed,
                          'executed_over_nominal': executed / nominal, 'dealt_by': m['stats'].get('deal'),
                          'value_executed': updates * (executed / nominal) / (m['dev_ms'] * 1e-3)}
    if m['e2e_ms']:
        n_local = wl.B // world
        out['e2e'] = {'value': updates / (m['e2e_ms'] * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': m['e2e_ms'],
                      'h2d_bytes_per_step': int(wl.data.nbytes + n_local * 2 * (8 + 4 + 16) + 2 * G * 8 + n_local * 16),
                      'd2h_bytes_per_step': int(T * G * 8 + T * 8 * (1 + len(wl.shape)) + wl.B * 16)}
        out['log_evidence'] = m['logE']
    return out


def measure_online(wl, bl, eng, torch, td, world, local_rank, clocks=False):
    """C5: T step() calls of an OnlineStudy through the public API (each step: H2D of the new segment, one batched
    launch over this rank's hypotheses, all-gather + D2H of the evidence increments), after a lead-in.  Device time =
    CUDA events around the whole window on the launching stream; per-kernel time from event brackets per call."""
    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    S = wl.study(bl)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        for d in wl.data[:wl.lead + 1]:
            S.step(d)
    timers = Timers(torch)
    plain_run = eng.run
    eng.run = timers.wrap(lambda which, *a, **kw: which, plain_run, name_of=eng.last_kernel)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank) if clocks else contextlib.nullcontext()
    try:
        with sampler as clk, contextlib.redirect_stdout(sink):
            barrier()
            timers.on = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for d in wl.data[wl.lead + 1:wl.lead + 1 + wl.T]:
                S.step(d)
            e1.record()
            barrier()
            wall_ms = 1e3 * (time.perf_counter() - t0)
            timers.on = False
    finally:
        eng.run = plain_run
    dev_ms = e0.elapsed_time(e1)
    ms, names = timers.collect()
    launches = eng.launch_count() - launches0
    times = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        td.all_reduce(times, op=td.ReduceOp.MAX)
    dev_ms, wall_ms = [float(x) for x in times.cpu()]
    post = S.marginalizedPosterior  # collective read
    return dict(dev_ms=dev_ms, wall_ms=wall_ms, ms=ms, names=names, launches=int(launches), logE=float(S.logEvidence),
                n_local=int(S._dev['Hr']), mean_rho=float(np.sum(post * S.grid[0])),
                clocks=clk.summary() if clocks else None)


def online_report(wl, m, peak, peak_src):
    T, G = wl.T, wl.G
    updates = float(wl.B) * T * G
    fwd = m['ms'].get('forward', [])
    step_kernel_ms = float(np.mean(fwd)) if fwd else float('nan')
    # algorithmic bytes per cell and step: K7 reads the state and writes the unnormalised cells, K8 reads them and writes
    # the normalised state: 32 B
    gbs = 32.0 * m['n_local'] * G / (step_kernel_ms * 1e-3) / 1e9
    name = m['names'].get('forward', 'online2d') + '_kernels'
    traffic, tag = ncu_traffic(wl.key, 'online2d_tile_kernel')
    roofline = {'bound': 'hbm', 'kernel': name, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
                'traffic': traffic, 'traffic_source': ('committed ncu capture profiles/%s (tile kernel only)' % tag) if tag else None,
                'peak_source': peak_src, 'kernels': {name: {'ms': step_kernel_ms, 'GBps': gbs, 'frac': gbs / peak,
                                                            'bytes_per_cell': 32.0}},
                'kernel_share_of_step': step_kernel_ms * len(fwd) / m['dev_ms'] if fwd else None}
    return {'value': updates / (m['dev_ms'] * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': m['dev_ms'],
            'ms_per_online_step': m['dev_ms'] / T, 'gpu_launches_per_step': m['launches'], 'roofline': roofline,
            'e2e': {'value': updates / (m['wall_ms'] * 1e-3), 'unit': 'cell-updates/s', 'ms_per_step': m['wall_ms'],
                    'ms_per_online_step': m['wall_ms'] / T,
                    'h2d_bytes_per_step': int(T * 2 * 8), 'd2h_bytes_per_step': int(T * wl.B * 8)},
            'log_evidence': m['logE'], 'posterior_mean_rho': m['mean_rho']}


# ------------------------------------------------------------------------------------------------ CPU port
def cpu_port_sample(key, args_dict, world, share, T_cpu):
    """Reference-style NumPy loop (oracle/np_oracle.py) on a share of the workload's sweep -- share = (i, n, per):
    chunk i of n of `n * per` combos spread evenly over the sweep -- and its first T_cpu data points.  Returns
    (cell updates, seconds, combos fitted, combos of the sweep).  The lowering comes from the product's host code;
    the arithmetic timed here is NumPy/SciPy only."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import bayesloop_b200 as bl
    import helpers
    import np_oracle
    wl = WORKLOADS[key](argparse.Namespace(**args_dict), world)

    def pick(B):
        i, n, per = share
        rows_all = np.unique(np.linspace(0, B - 1, n * per).round().astype(int))
        return [int(r) for r in np.array_split(rows_all, n)[i]]

    if wl.kind == 'online':
        S = wl.study(bl)
        S.rawData = np.asarray(wl.data[:2])
        pb = helpers.np_problem_online(S, wl.data[:T_cpu + 1])
        ops = helpers.online_ops(S)
        rows = pick(S.tmCount)
        t0 = time.perf_counter()
        np_oracle.online_steps(pb, ops, rows)
        return float(len(rows)) * pb.T * int(np.prod(pb.shape)), time.perf_counter() - t0, len(rows), S.tmCount
    S = wl.study(bl, T=T_cpu)
    if hasattr(S, '_prepareChangepoints'):
        S._prepareChangepoints(silent=True)
        ctx = S._lower(np.asarray(S.hyperGridValues, dtype=float), S.formattedTimestamps)
        ops, hp = ctx.ops, np.asarray(S.flatHyperPriorValues, dtype=float)
    else:
        ops, hp, _ = helpers.lowered(S)
    pb = helpers.np_problem(S)
    rows = pick(len(hp))
    t0 = time.perf_counter()
    np_oracle.hyper_fit(pb, ops, hp, rows=rows)
    return np_oracle.cell_updates(pb, len(rows)), time.perf_counter() - t0, len(rows), len(hp)


def _cpu_worker(job):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    with contextlib.redirect_stdout(io.StringIO()):
        return cpu_port_sample(*job)


def _cpu_worker_warm(_):
    """Imports of a worker process (torch, the package, the port), kept out of the timed steps whatever --warmup is."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    with contextlib.redirect_stdout(io.StringIO()):
        import bayesloop_b200  # noqa: F401
        import helpers  # noqa: F401
        import np_oracle  # noqa: F401
    time.sleep(0.2)  # long enough for every worker of the pool to take one of these
    return 0


CPU_SAMPLE = {'c2': (24, 4000), 'c3': (6, 40), 'c4': (6, 40), 'c5': (8, 12)}  # (combos, time steps) of the 1-core sample


def cpu_baseline(wl, args, world):
    n, T_cpu = CPU_SAMPLE[wl.key]
    T_cpu = min(T_cpu, wl.T)
    with contextlib.redirect_stdout(io.StringIO()):
        upd, sec, fitted, total = cpu_port_sample(wl.key, vars(args), world, (0, 1, n), T_cpu)
    return {'value': upd / sec, 'unit': 'cell-updates/s', 'cores': 1, 'kind': 'port',
            'sample': '%d of %d combos (evenly spaced), first %d of %d time steps, oracle/np_oracle.py '
                      '(NumPy + scipy.ndimage.gaussian_filter1d), %.1f s' % (fitted, total, T_cpu, wl.T, sec)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (NumPy port; the reference itself is Python and is not installed on
    the GPU box) on all host cores, rows of the sweep split across worker processes like HyperStudy._parallelFit
    (core.py:1463-1465).  Each step is a bounded sample of the workload."""
    if rank != 0:
        return
    import multiprocessing as mp
    wl = WORKLOADS[args.config](args, args.gpus)
    cores = max(1, (os.cpu_count() or 1))
    workers = min(cores, 64)
    per_worker = {'c2': 4, 'c3': 2, 'c4': 2, 'c5': 2}[wl.key]
    T_cpu = min({'c2': args.cpu_T, 'c3': 24, 'c4': 24, 'c5': 8}[wl.key], wl.T)
    jobs = [(wl.key, vars(args), args.gpus, (i, workers, per_worker), T_cpu) for i in range(workers)]
    ctx = mp.get_context('spawn')
    times = []
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker_warm, range(4 * workers), chunksize=1)
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    updates = float(sum(r[0] for r in res))
    ms = 1e3 * float(np.mean(times))
    value = updates / (ms / 1e3)
    sample = '%d of %d combos (evenly spaced), first %d of %d time steps, %d worker processes' % (
        sum(r[2] for r in res), res[0][3], T_cpu, wl.T, workers)
    line = {
        'impl': 'reference', 'metric': 'grid_cell_updates_per_s', 'value': value, 'unit': 'cell-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': wl.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': wl.describe(),
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': workers, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--T', type=int, default=0, help='time steps (default: 10000 for c2, window of 200 / 400 / 200 for c3 / c4 / c5)')
    ap.add_argument('--grid', type=int, default=1000, help='c2: cells of the 1-D grid')
    ap.add_argument('--sigma-max', dest='sigma_max', type=float, default=0.2)
    ap.add_argument('--combos-per-gpu', dest='combos', type=int, default=512, help='c2: sigma values per GPU')
    ap.add_argument('--hyper', type=int, default=64, help='c3: values per hyper-parameter axis')
    ap.add_argument('--changepoints', type=int, default=100, help='c4: change-point values')
    ap.add_argument('--cpu-T', dest='cpu_T', type=int, default=4000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', dest='no_extra', action='store_true', help='headline only: skip the other configs')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3  # timing rule: at least 3 warm-up steps

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
    peak, peak_src = hbm_peak()

    def run_config(key, steps, warmup, headline):
        wl = WORKLOADS[key](args if (headline or not args.T) else argparse.Namespace(**dict(vars(args), T=0)), world)
        if wl.kind == 'online':
            m = measure_online(wl, bl, eng, torch, td, world, local, clocks=headline)
            rep = online_report(wl, m, peak, peak_src)
        else:
            m = measure_sweep(wl, bl, eng, torch, td, world, steps, warmup, local, e2e=True, clocks=headline)
            rep = sweep_report(wl, m, peak, peak_src, world)
        rep['config'] = wl.describe()
        rep['scaling'] = wl.scaling
        rep['n_gpus'] = world
        return wl, m, rep

    wl, m, rep = run_config(args.config, args.steps, args.warmup, True)
    line = None
    if rank == 0:
        line = {'metric': 'grid_cell_updates_per_s', 'value': rep['value'], 'unit': 'cell-updates/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': rep['ms_per_step'], 'higher_is_better': True,
                'scaling': wl.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': rep['config'],
                'clocks': m['clocks'], 'e2e': rep['e2e'], **({'sharing': rep['sharing']} if 'sharing' in rep else {}),
                'gpu_launches': rep['gpu_launches_per_step'] * (args.steps if wl.kind == 'sweep' else 1),
                'roofline': rep['roofline'], 'log_evidence': rep.get('log_evidence')}
        if wl.kind == 'online':
            line['ms_per_online_step'] = rep['ms_per_online_step']
    extra = {}
    if args.config == 'c2' and not args.no_extra:
        # the other BASELINE configs, measured at EVERY N (strong scaling); never lose the headline over them
        if world == 1:
            try:  # narrow sweep of SURVEY.md 8d (sigma <= 0.05, radius <= 17): the HBM-leaning regime of C2
                narrow = C2(argparse.Namespace(**dict(vars(args), sigma_max=0.05)), world)
                mn = measure_sweep(narrow, bl, eng, torch, td, world, 3, 2, local, e2e=False)
                rn = sweep_report(narrow, mn, peak, peak_src, world)
                extra['c2_narrow'] = {'workload': 'C2 with the narrow sweep cint(0,0.05,%d): kernel radius <= 17' % narrow.B,
                                      'value': rn['value'], 'unit': 'cell-updates/s', 'ms_per_step': rn['ms_per_step'],
                                      'kernels': {k: {'ms': v['ms'], 'hbm_frac': v['frac']}
                                                  for k, v in rn['roofline']['kernels'].items()}}
            except Exception as exc:
                extra['c2_narrow'] = {'error': repr(exc)}
        for key in ('c3', 'c4', 'c5'):
            try:
                wl2, m2, rep2 = run_config(key, 2, 1, False)
                if rank == 0 and world == 1 and not args.no_cpu_baseline and key != 'c4':
                    rep2['cpu_baseline'] = cpu_baseline(wl2, args, world)
                    rep2['vs_cpu_1core'] = rep2['value'] / rep2['cpu_baseline']['value']
                extra[key] = rep2
            except Exception as exc:  # identical control flow on all ranks: the failure is deterministic
                extra[key] = {'error': repr(exc)}
                torch.cuda.empty_cache()
    if rank == 0:
        if extra:
            line['extra'] = extra
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(wl, args, world)
        print(json.dumps(line), flush=True)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == '__main__':
    main()
