#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2i_pytest_gpu.log
bash tools/sanitize.sh
timeout 300 python bench.py --steps 3 --no-extra --no-cpu-baseline 2> gpurun_out/r2i_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C2 ms/step %.1f e2e %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step']), {k: round(v['ms'], 2) for k, v in d['roofline']['kernels'].items()})"
