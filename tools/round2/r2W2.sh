#!/bin/bash
# where the end-to-end time of a C2 fit goes (host side)
mkdir -p gpurun_out
timeout 200 python tools/e2e_breakdown.py 6 > gpurun_out/r2W2_e2e_breakdown.txt 2>&1
tail -14 gpurun_out/r2W2_e2e_breakdown.txt | cut -c1-400
