#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c5_size" > gpurun_out/r2e_c5test.log 2>&1; grep -E "Error|assert|passed|failed" gpurun_out/r2e_c5test.log | head -20
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2e_online_launches.csv python tools/exp_online.py 512 14 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2e_online_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik][:60]].append(float(r[iv].replace(',', '')))
for k, v in d.items():
    print(k, len(v), 'mean us %.1f' % (sum(v[len(v)//2:]) / len(v[len(v)//2:]) / 1e3))
PY
timeout 300 python bench.py --steps 3 --no-extra --no-cpu-baseline 2> gpurun_out/r2e_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 ms/step %.1f e2e %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v['ms'], 2) for k, v in d['roofline']['kernels'].items()})"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2e_pytest_gpu.log
