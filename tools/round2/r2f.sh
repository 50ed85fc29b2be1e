#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "online or stream_kernels" 2>&1 | tail -3
timeout 90 python tools/exp_online.py 512 100 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_online_launches.csv python tools/exp_online.py 512 14 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2f_online_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik][:60]].append(float(r[iv].replace(',', '')))
for k, v in d.items():
    print(k, len(v), 'mean us %.1f' % (sum(v[len(v)//2:]) / len(v[len(v)//2:]) / 1e3))
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:online2d_tile -s 6 -c 1 -f -o gpurun_out/r2f_online2d_tile python tools/exp_online.py 512 14 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
