#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline 2> gpurun_out/r2r_c4.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('C4 ms/step %.1f value %.3g e2e %.3g' % (d['ms_per_step'], d['value'], d['e2e']['value']), {k: (round(v['ms'], 1), v['launches_per_step']) for k, v in d['roofline']['kernels'].items()}, 'kernel share', d['roofline']['kernel_share_of_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2r_c4_launches.csv python bench.py --config c4 --changepoints 20 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2r_c4_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik][:70]].append(float(r[iv].replace(',', '')))
tot = sum(sum(v) for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1]))[:14]:
    print('%-72s n=%5d total ms %9.1f share %.3f' % (k, len(v), sum(v) / 1e6, sum(v) / tot))
PY
