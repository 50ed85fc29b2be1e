#!/bin/bash
# the full GPU suite on the last commit of the round
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/r2last_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2last_pytest_gpu.log
