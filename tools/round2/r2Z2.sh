#!/bin/bash
# DMMA groups starting at offset -R (half a group fewer per convolution): quick bench first, then the suite
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "logE", d.get("log_evidence"))'
timeout 120 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2Z2.err | tee gpurun_out/r2Z2.json | python -c "$summ" "groups from -R" || { echo "bench failed or hung"; tail -3 gpurun_out/r2Z2.err; exit 1; }
timeout 400 python -m pytest tests -m gpu -q -x -k "not multi" > gpurun_out/r2Z2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2Z2_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2Z2_pytest_gpu.log | head -20
