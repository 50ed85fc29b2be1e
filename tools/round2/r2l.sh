#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "online" 2>&1 | tail -2
timeout 90 python tools/exp_online.py 512 100 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2l_online_launches.csv python tools/exp_online.py 512 14 > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2l_online_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik][:60]].append(float(r[iv].replace(',', '')))
for k, v in d.items():
    print(k, len(v), 'mean us %.1f' % (sum(v[len(v)//2:]) / len(v[len(v)//2:]) / 1e3))
PY
timeout 120 python tools/e2e_breakdown.py 4 2>&1 | tail -8
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline 2> gpurun_out/r2l_c4.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('C4 ms/step %.1f value %.3g e2e %.3g' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
timeout 600 python bench.py --config c3 --steps 2 --no-cpu-baseline 2> gpurun_out/r2l_c3.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('C3 ms/step %.1f value %.3g e2e %.3g' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
