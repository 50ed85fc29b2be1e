#!/bin/bash
# DMMA 1-D kernels: natural layout, one 16-byte fragment load per two matrix instructions (k-groups of 8)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not multi" > gpurun_out/r2A_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2A_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2A_pytest_gpu.log | head -30
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2A_mma.err | tee gpurun_out/r2A_mma.json | python -c "$summ" mma
BLG_TRACE=gpurun_out/r2A_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
python tools/sm_timeline.py gpurun_out/r2A_trace | grep -v "SM  "
