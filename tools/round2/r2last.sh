#!/bin/bash
# last sanity check of the round on the final library: smoke() and the headline bench line
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2last.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('C2 ms/step %.1f value %.4g e2e %.4g (%.1f ms)' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v['ms'], 2) for k, v in d['roofline']['kernels'].items()}, 'logE', d.get('log_evidence'))"
