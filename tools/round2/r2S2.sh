#!/bin/bash
# both DMMA kernels: short store path (warp-uniform), kappa in front of the convolution, sequential backward epilogue without spills
# trace as a template parameter): parity, bench, per-step event trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not multi" > gpurun_out/r2S2_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2S2_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2S2_pytest_gpu.log | head -20
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "logE", d.get("log_evidence"))'
timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2S2.err | tee gpurun_out/r2S2.json | python -c "$summ" "lean loops v2"
