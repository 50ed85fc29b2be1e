#!/bin/bash
# per-step event trace of the forward DMMA kernel (32 steps in the middle of a T = 2000 sweep)
mkdir -p gpurun_out
BLG_TRACE=gpurun_out/r2E_trace timeout 200 python tools/trace_c2.py 2000 2>&1 | tail -3
ls -la gpurun_out | tail -6
