#!/bin/bash
# ncu source-level capture of the DMMA 1-D kernels at T = 2000
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fast1d_mma' --launch-skip 6 -c 2 \
    -f -o gpurun_out/r2z_c2_mma python bench.py --T 2000 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2z_ncu_c2.log 2>&1
tail -2 gpurun_out/r2z_ncu_c2.log
python tools/ncu_regions.py gpurun_out/r2z_c2_mma.ncu-rep fast1d_mma 25 > gpurun_out/r2z_regions.txt 2>&1
head -40 gpurun_out/r2z_regions.txt
ncu -i gpurun_out/r2z_c2_mma.ncu-rep --page raw --csv > gpurun_out/r2z_raw.csv 2>/dev/null
