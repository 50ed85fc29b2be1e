#!/bin/bash
# diagnostic: forward DMMA kernel with the convolution / the FP64 part of the epilogue removed (results are garbage)
mkdir -p gpurun_out
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f" % d["ms_per_step"], {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})'
cd bayesloop_b200/csrc; cp libblgrid.so /tmp/libblgrid_base.so; cd ../..
for v in NOCONV NOEPI; do
  cp bayesloop_b200/csrc/libblgrid_exp_$v.so.x bayesloop_b200/csrc/libblgrid.so
  timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-extra 2> gpurun_out/r2D_$v.err | tee gpurun_out/r2D_$v.json | python -c "$summ" "$v"
done
cp /tmp/libblgrid_base.so bayesloop_b200/csrc/libblgrid.so
ncu --set full --clock-control none --import-source on -k regex:'fwd_fast1d_mma' --launch-skip 3 -c 1 \
    -f -o gpurun_out/r2D_fwd python bench.py --T 2000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2D_ncu.log 2>&1
ncu -i gpurun_out/r2D_fwd.ncu-rep --page raw --csv > gpurun_out/r2D_raw.csv 2>/dev/null
ncu -i gpurun_out/r2D_fwd.ncu-rep --page source --csv > gpurun_out/r2D_src.csv 2>/dev/null
rm -f gpurun_out/r2D_fwd.ncu-rep
ls -la gpurun_out | tail -5
