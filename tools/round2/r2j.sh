#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_gpu.log 2>&1
tail -6 gpurun_out/r2j_pytest_gpu.log
grep -E "^E  " gpurun_out/r2j_pytest_gpu.log | head -20
timeout 600 python bench.py --config c4 --steps 2 --no-cpu-baseline > gpurun_out/r2j_c4.json 2> gpurun_out/r2j_c4.err
tail -3 gpurun_out/r2j_c4.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2j_c4.json"))
print("C4 ms/step %.1f value %.3g e2e %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("sharing"),
      {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})
PY
