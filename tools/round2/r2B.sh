#!/bin/bash
# DMMA kernels + pacing of the chains of an SM
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
for cfg in "4 8" "2 4" "8 16" "4 16" "16 64"; do
  set -- $cfg
  BLG_WS_PACE_EVERY=$1 BLG_WS_PACE_SKEW=$2 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2B_pace_$1_$2.err | tee gpurun_out/r2B_pace_$1_$2.json | python -c "$summ" "pace every $1 skew $2"
done
BLG_WS_PACE_EVERY=4 BLG_WS_PACE_SKEW=8 BLG_TRACE=gpurun_out/r2B_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
python tools/sm_timeline.py gpurun_out/r2B_trace | grep -v "SM  "
