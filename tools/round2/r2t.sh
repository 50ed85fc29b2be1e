#!/bin/bash
# ws kernels: exact tap counts + uneven split (3 warps x 9 + 1 warp x 5 cells at G = 1000) against the even split
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "not multi" > gpurun_out/r2t_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2t_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2t_pytest_gpu.log | head
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra 2> gpurun_out/r2t_uneven.err | tee gpurun_out/r2t_uneven.json | python -c "$summ" uneven
BLG_WS_EVEN=1 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extra 2> gpurun_out/r2t_even.err | tee gpurun_out/r2t_even.json | python -c "$summ" even
BLG_TRACE=gpurun_out/r2t_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
