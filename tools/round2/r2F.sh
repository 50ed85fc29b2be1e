#!/bin/bash
# diagnostic: both DMMA kernels with the FP64 arithmetic of the epilogue removed (results are garbage)
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f" % d["ms_per_step"], {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
cp bayesloop_b200/csrc/libblgrid.so /tmp/libblgrid_base.so
cp bayesloop_b200/csrc/libblgrid_exp_NOEPI.so.x bayesloop_b200/csrc/libblgrid.so
timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra 2> gpurun_out/r2F_NOEPI.err | tee gpurun_out/r2F_NOEPI.json | python -c "$summ" "NOEPI fwd+bwd"
cp /tmp/libblgrid_base.so bayesloop_b200/csrc/libblgrid.so
