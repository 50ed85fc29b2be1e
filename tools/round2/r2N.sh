#!/bin/bash
# interleaved 1-D kernels: mbarrier waits between the compute warps instead of spinning on the counters
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "interleaved" > gpurun_out/r2N_pytest_il.log 2>&1
tail -3 gpurun_out/r2N_pytest_il.log | cut -c1-300
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
BLG_IL=1 timeout 200 python bench.py --steps 3 --no-cpu-baseline --no-extra 2> gpurun_out/r2N_il.err | tee gpurun_out/r2N_il.json | python -c "$summ" "interleaved"
tail -3 gpurun_out/r2N_il.err
