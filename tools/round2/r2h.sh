#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2h_pytest_gpu.log
timeout 90 python tools/exp_online.py 512 100 2>&1 | tail -2
bash tools/sanitize.sh
