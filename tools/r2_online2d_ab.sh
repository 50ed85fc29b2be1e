#!/bin/bash
# First GPU call of round 2 for the tiled online step (K7): parity tests and the C5-shaped timing under each of the
# variants that were written after the last GPU run of round 1 (cp.async tile loads, 32-row tiles with two CTAs per SM).
#   gpurun --timeout 300 -- 'bash tools/r2_online2d_ab.sh'
mkdir -p gpurun_out
# random operator programs, CUDA against the CPU oracle (tests/test_gpu_fuzz.py)
BLG_TEST_FUZZ=150 timeout 240 python -m pytest tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -3
# cases added after the last GPU minute of round 1 (tests/cases.py: GPU_DEFERRED)
BLG_TEST_DEFERRED=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sympy or scipy or numpy" 2>&1 | tail -3
for v in "BLG_ONLINE2D=0" "BLG_ONLINE2D=1" "BLG_ONLINE2D_ASYNC=1" "BLG_ONLINE2D_TH=32" "BLG_ONLINE2D_ASYNC=1 BLG_ONLINE2D_TH=32"; do
    tag=$(echo "$v" | tr ' =' '__')
    echo "== $v"
    if [ "$v" != "BLG_ONLINE2D=0" ]; then
        env $v timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -q -k online2d 2>&1 | tail -1
    fi
    env $v timeout 60 python tools/exp_online.py 512 60 2>&1 | tail -2 | tee gpurun_out/r2_online_$tag.log
done
# note: tests/test_gpu_parity.py sets BLG_ONLINE2D=1 itself; the other variables pass through to the library
# one ncu capture of the tile kernel under the most aggressive variant (compare with profiles/r1k_ncu_online2d.json)
BLG_ONLINE2D_ASYNC=1 BLG_ONLINE2D_TH=32 timeout 120 ncu --set full --clock-control none --import-source on -k regex:online2d_tile \
    -s 20 -c 1 -f -o gpurun_out/r2_online2d_tile_async_th32 python tools/exp_online.py 512 30 > gpurun_out/r2_ncu_online2d.log 2>&1
tail -2 gpurun_out/r2_ncu_online2d.log
