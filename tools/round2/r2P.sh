#!/bin/bash
# forward DMMA kernel with the deferred sum (one FP64 level on the critical path of a step) + one-Newton reciprocal in the
# backward local-evidence sum: full GPU suite, bench against the undeferred protocol, per-step event trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not multi" > gpurun_out/r2P_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2P_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2P_pytest_gpu.log | head -20
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "logE", d.get("log_evidence"))'
timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2P_defer.err | tee gpurun_out/r2P_defer.json | python -c "$summ" "deferred sum"
BLG_NO_MMA_DEFER=1 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2P_nodefer.err | tee gpurun_out/r2P_nodefer.json | python -c "$summ" "sum inside the step"
BLG_TRACE=gpurun_out/r2P_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
python tools/sm_timeline.py gpurun_out/r2P_trace 2>/dev/null | grep -v "SM  " | head -8
