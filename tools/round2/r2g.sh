#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "online" > gpurun_out/r2g_online_tests.log 2>&1
grep -E "^E |passed|failed|^FAILED" gpurun_out/r2g_online_tests.log | head -40
