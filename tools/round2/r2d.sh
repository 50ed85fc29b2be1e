#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "online" 2>&1 | tail -3
for v in "BLG_ONLINE2D_ASYNC=1" "BLG_ONLINE2D_ASYNC=0"; do
  echo "== $v"; env $v timeout 90 python tools/exp_online.py 512 100 2>&1 | tail -2
done
BLG_TRACE=gpurun_out/r2d_ws11 timeout 120 python tools/trace_c2.py 2000
BLG_WS_M=9 BLG_WS_NT=160 BLG_TRACE=gpurun_out/r2d_ws9 timeout 120 python tools/trace_c2.py 2000
for v in "9 160"; do
  set -- $v
  echo "== ws M=$1 NT=$2"
  BLG_WS_M=$1 BLG_WS_NT=$2 timeout 300 python bench.py --steps 3 --no-extra --no-cpu-baseline 2> gpurun_out/r2d_ws_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.1f' % d['ms_per_step'], {k: round(v['ms'], 2) for k, v in d['roofline']['kernels'].items()})"
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2d_pytest_gpu.log
ls gpurun_out | grep r2d
