#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2q_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2q_pytest_gpu.log | head
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline 2> gpurun_out/r2q_c4.err | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][0]; print('C4 ms/step %.1f value %.3g e2e %.3g' % (d['ms_per_step'], d['value'], d['e2e']['value']), {k: round(v['ms'], 1) for k, v in d['roofline']['kernels'].items()})"
timeout 90 python tools/exp_online.py 512 100 2>&1 | tail -2
