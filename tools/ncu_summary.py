"""Turn the ncu artefacts of tools/profile_round.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:

    python tools/ncu_summary.py r2m

  profiles/<tag>_launches_bench.csv   per-kernel launch list of the bench command (count, total / mean duration, share)
  profiles/<tag>_ncu_summary.json     key `ncu --set full` metrics per kernel (read by bench.py for roofline.traffic)
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

METRICS = {
    'gpu__time_duration.sum': 'duration_ms',
    'dram__bytes_read.sum': 'dram_bytes_read',
    'dram__bytes_write.sum': 'dram_bytes_write',
    'lts__t_bytes.sum': 'l2_bytes',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct_of_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed': 'fp64_pipe_pct_of_elapsed',
    # mma.m8n8k4.f64 (DMMA) is counted on the tensor pipe, not on the FP64 pipe
    'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active': 'dmma_pipe_pct_of_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed': 'smem_pipe_pct',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__cluster_size': 'cluster_size',
    'launch__cluster_max_active': 'clusters_resident',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem_per_block',
    'smsp__inst_executed.sum': 'warp_instructions',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_scoreboard',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_scoreboard',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_pipe',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio': 'stall_membar',
}
UNIT_SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'second': 1e3,
              'msecond': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6}


def raw_page(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        entry = {'kernel': re.sub(r'^void |\(.*$', '', r[hdr.index('Kernel Name')]).strip()}
        for i, h in enumerate(hdr):
            if h in METRICS and r[i] != '':
                val = float(r[i].replace(',', ''))
                entry[METRICS[h]] = val * UNIT_SCALE.get(units[i], 1.0)
        res.append(entry)
    return res


def launches(tag):
    path = os.path.join(OUT, tag + '_launches_bench.csv')
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = {}
    for r in rows[1:]:
        name = re.sub(r'^void |\(.*$', '', r[ik]).strip()
        ms = float(r[iv].replace(',', '')) * UNIT_SCALE.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    dst = os.path.join(PROF, tag + '_launches_bench.csv')
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 '
                '--no-cpu-baseline --no-extra (cold-cache, serialised launches: shares matter, not absolutes)\n')
        f.write('kernel,launches,total_ms,mean_ms,share\n')
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%s,%d,%.3f,%.3f,%.4f\n' % (name, n, ms, ms / n, ms / total))
    return dst


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
    os.makedirs(PROF, exist_ok=True)
    print(launches(tag))
    summary = {'tag': tag, 'command': 'tools/profile_round.sh ' + tag,
               'note': 'ncu --set full --clock-control none; bytes and durations are per launch'}
    for key, rep, what in (('bench_c2', '_c2_ws', 'C2 at the bench shape (512 combos, T = 10000)'),
                           ('bench_c3', '_c3_cluster', 'C3 grid and hyper-ranges, 8 x 8 of the hyper-grid, 200 steps'),
                           ('bench_c5', '_c5_online', 'C5 at the full shape (512 x 512, 256 hypotheses), one step')):
        path = os.path.join(OUT, tag + rep + '.ncu-rep')
        if os.path.exists(path):
            summary[key] = {e['kernel'].split('<')[0].replace('blg::', ''): e for e in raw_page(path)}
            summary[key + '_workload'] = what
    dst = os.path.join(PROF, tag + '_ncu_summary.json')
    with open(dst, 'w') as f:
        json.dump(summary, f, indent=1, sort_keys=True)
    print(dst)


if __name__ == '__main__':
    main()
