#!/bin/bash
# final kernels: pacing variants (default every 4 / skew 8)
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f" % d["ms_per_step"], {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
for cfg in "4 8" "4 0" "8 16" "2 6"; do
  set -- $cfg
  BLG_WS_PACE_EVERY=$1 BLG_WS_PACE_SKEW=$2 timeout 100 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2pace_$1_$2.err | python -c "$summ" "pace every $1 skew $2"
done
