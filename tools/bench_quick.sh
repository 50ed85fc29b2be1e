#!/bin/bash
# quick device-side numbers of the C2 sweep (no CPU baseline): tools/bench_quick.sh <tag>
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
tail -3 gpurun_out/bench_$1.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$1.json"))
print("$1", "ms/step %.1f" % d["ms_per_step"], "value %.3g" % d["value"], "e2e ms %.1f" % d["e2e"]["ms_per_step"], {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, d["log_evidence"])
PY
