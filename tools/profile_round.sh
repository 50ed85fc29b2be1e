#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun): launch list of the bench command, full captures of
# the dominant kernels.  Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.
set -x
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# C2 (headline): forward + backward + accumulate of one timed sweep at the bench shape (skip the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:'fast1d_ws|accumulate' --launch-skip 9 -c 3 \
    -o gpurun_out/${TAG}_c2_ws python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-c3 > gpurun_out/${TAG}_ncu_c2.log 2>&1
# C3 sample: cluster-resident 2-D kernels
ncu --set full --clock-control none --import-source on -k regex:cluster2d --launch-skip 2 -c 2 \
    -o gpurun_out/${TAG}_c3_cluster python tools/exp_2d.py 256 200 6 0.1 > gpurun_out/${TAG}_ncu_c3.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_c2.log; tail -n 2 gpurun_out/${TAG}_ncu_c3.log
