#!/bin/bash
# re-entry check of HEAD: full GPU suite, C4 with the batched share_apply, DMMA/DFMA concurrency probe, per-SM trace
mkdir -p gpurun_out
timeout 60 tools/micro/dmma_mix_bench > gpurun_out/r2s_dmma_mix.txt 2>&1; cat gpurun_out/r2s_dmma_mix.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2s_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2s_pytest_gpu.log | head
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline 2> gpurun_out/r2s_c4.err > gpurun_out/r2s_c4.json
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2s_c4.json') if l.startswith('{')][0]; print('C4 ms/step %.1f value %.3g e2e %.3g' % (d['ms_per_step'], d['value'], d['e2e']['value']), {k: (round(v['ms'], 1), v['launches_per_step']) for k, v in d['roofline']['kernels'].items()}, 'kernel share', d['roofline']['kernel_share_of_step'])"
BLG_TRACE=gpurun_out/r2s_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
ls gpurun_out | head -30
