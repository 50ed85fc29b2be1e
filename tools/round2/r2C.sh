#!/bin/bash
# DMMA kernels: 8 compute warps x 2 tiles (288 threads, 56 registers, 4 chains per SM) against 4 x 4
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
BLG_MMA_NT=288 timeout 300 python -m pytest tests -m gpu -q -x -k "poisson or golden" > gpurun_out/r2C_pytest.log 2>&1; tail -3 gpurun_out/r2C_pytest.log
BLG_MMA_NT=288 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2C_288.err | tee gpurun_out/r2C_288.json | python -c "$summ" "8x2 tiles"
BLG_MMA_NT=288 BLG_WS_PACE_EVERY=4 BLG_WS_PACE_SKEW=8 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2C_288p.err | tee gpurun_out/r2C_288p.json | python -c "$summ" "8x2 tiles paced"
timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2C_160.err | tee gpurun_out/r2C_160.json | python -c "$summ" "4x4 tiles"
BLG_MMA_NT=288 BLG_TRACE=gpurun_out/r2C_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
python tools/sm_timeline.py gpurun_out/r2C_trace | grep -v "SM  "
