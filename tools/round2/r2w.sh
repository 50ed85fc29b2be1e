#!/bin/bash
# DMMA variant of the 1-D kernels: record the kernel families, parity run with them, A/B against the DFMA kernels
mkdir -p gpurun_out
rm -f gpurun_out/kernel_families.json
BLG_RECORD_FAMILIES=gpurun_out/kernel_families.json timeout 600 python -m pytest tests -m gpu -q -k "golden" > gpurun_out/r2w_pytest_record.log 2>&1
tail -3 gpurun_out/r2w_pytest_record.log
cp gpurun_out/kernel_families.json tests/golden/kernel_families.json
sed -i "s/'poisson_c2_small': 'fast1d_ws', 'poisson_wide_kernels': 'fast1d_ws'/'poisson_c2_small': 'fast1d_mma', 'poisson_wide_kernels': 'fast1d_mma'/" tests/test_gpu_parity.py
timeout 900 python -m pytest tests -m gpu -q -k "not multi" > gpurun_out/r2w_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2w_pytest_gpu.log; grep -E "^E  |^FAILED" gpurun_out/r2w_pytest_gpu.log | head -30
summ='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][0]; print(sys.argv[1], "ms/step %.1f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "logE", d.get("extra", {}).get("log_evidence"))'
timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2w_mma.err | tee gpurun_out/r2w_mma.json | python -c "$summ" mma
BLG_NO_MMA=1 timeout 300 python bench.py --steps 4 --no-cpu-baseline --no-extra 2> gpurun_out/r2w_ws.err | tee gpurun_out/r2w_ws.json | python -c "$summ" ws
tail -3 gpurun_out/r2w_mma.err
