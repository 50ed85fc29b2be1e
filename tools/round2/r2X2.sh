#!/bin/bash
# e2e after caching cudaMemGetInfo in the engine
mkdir -p gpurun_out
timeout 200 python tools/e2e_breakdown.py 6 > gpurun_out/r2X2_e2e_breakdown.txt 2>&1
grep "^fit" gpurun_out/r2X2_e2e_breakdown.txt
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g (%.1f ms)" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
timeout 150 python bench.py --steps 5 --no-cpu-baseline --no-extra 2> gpurun_out/r2X2.err | tee gpurun_out/r2X2.json | python -c "$summ" "bench"
