#!/bin/bash
# ncu evidence of one round (run on the GPU box through gpurun): launch list of the bench command, `--set full` captures
# of the dominant kernels of every config.  Outputs land in gpurun_out/; tools/ncu_summary.py <tag> turns them into
# profiles/<tag>_launches_bench.csv and profiles/<tag>_ncu_summary.json (read by bench.py for roofline.traffic).
#   gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2m'
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# C2 (headline): forward + backward + accumulate of one timed sweep at the bench shape (3 warm-up sweeps skipped)
ncu --set full --clock-control none --import-source on -k regex:'fast1d_il|fast1d_mma|fast1d_ws|accumulate_kernel' --launch-skip 9 -c 3 \
    -f -o gpurun_out/${TAG}_c2_ws python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu_c2.log 2>&1
# C3: cluster-resident 2-D kernels, 8 x 8 of the hyper-grid, window of 200 steps
ncu --set full --clock-control none --import-source on -k regex:cluster2d --launch-skip 6 -c 2 \
    -f -o gpurun_out/${TAG}_c3_cluster python bench.py --config c3 --hyper 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c3.log 2>&1
# C5: tiled online step at the full shape (512 x 512, 256 hypotheses)
ncu --set full --clock-control none --import-source on -k regex:'online2d_tile|online2d_finish' -s 12 -c 2 \
    -f -o gpurun_out/${TAG}_c5_online python tools/exp_online.py 512 14 > gpurun_out/${TAG}_ncu_c5.log 2>&1
tail -n 2 gpurun_out/${TAG}_ncu_c2.log gpurun_out/${TAG}_ncu_c3.log gpurun_out/${TAG}_ncu_c5.log
