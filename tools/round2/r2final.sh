#!/bin/bash
# final run of the round: closing artefacts (reference arm, default bench, GPU suite, smoke) + ncu evidence of the C2 kernels
mkdir -p gpurun_out
bash tools/round2/r2I.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2F2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2F2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fast1d_mma|accumulate_kernel' --launch-skip 9 -c 3 \
    -f -o gpurun_out/r2F2_c2_ws python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2F2_ncu_c2.log 2>&1
tail -n 2 gpurun_out/r2F2_ncu_c2.log
