#!/bin/bash
# ws kernels: pacing of the chains that share an SM (service warps hold a chain back while it is ahead of its peers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "not multi" > gpurun_out/r2u_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2u_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2u_pytest_gpu.log | head
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
for cfg in "4 8" "1 2" "1 1" "2 4" "8 16" "16 32" "4 0"; do
  set -- $cfg
  BLG_WS_PACE_EVERY=$1 BLG_WS_PACE_SKEW=$2 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2u_pace_$1_$2.err | tee gpurun_out/r2u_pace_$1_$2.json | python -c "$summ" "pace every $1 skew $2"
done
BLG_TRACE=gpurun_out/r2u_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
python tools/sm_timeline.py gpurun_out/r2u_trace
