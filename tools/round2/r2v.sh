#!/bin/bash
# ncu source-level capture of the ws kernels (uneven split, exact taps) at T = 2000
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fast1d_ws' --launch-skip 6 -c 2 \
    -f -o gpurun_out/r2v_c2_ws python bench.py --T 2000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2v_ncu_c2.log 2>&1
tail -2 gpurun_out/r2v_ncu_c2.log
python tools/ncu_regions.py gpurun_out/r2v_c2_ws.ncu-rep fast1d_ws 25 > gpurun_out/r2v_regions.txt 2>&1
head -80 gpurun_out/r2v_regions.txt
ncu -i gpurun_out/r2v_c2_ws.ncu-rep --page raw --csv > gpurun_out/r2v_raw.csv 2>/dev/null
ls -la gpurun_out/
