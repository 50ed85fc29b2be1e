#!/bin/bash
# Round-2 GPU call: full parity suite (kernel families recorded), smoke, headline bench.
#   gpurun --timeout 900 -- 'bash tools/round2/r2b_tests.sh'
mkdir -p gpurun_out
rm -f gpurun_out/kernel_families.json
BLG_RECORD_FAMILIES=gpurun_out/kernel_families.json timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2b_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python bench.py --steps 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2b_bench.json"))
print("bench ms/step %.1f value %.3g e2e %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]),
      {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "c3", d.get("extra", {}).get("c3_sample", {}).get("ms_per_step"))
PY
