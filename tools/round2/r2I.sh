#!/bin/bash
# closing run of the round: reference arm, default bench, full GPU suite, smoke
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2I_reference.json 2> gpurun_out/r2I_reference.err
tail -c 600 gpurun_out/r2I_reference.json
( time timeout 900 python bench.py ) > gpurun_out/r2I_bench.json 2> gpurun_out/r2I_bench.err
tail -4 gpurun_out/r2I_bench.err
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2I_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2I_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
