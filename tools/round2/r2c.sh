#!/bin/bash
# Round-2 GPU call C: parity after the lagged-scale rewrite of the ws kernels, new bench (all configs), ws variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2c_pytest_gpu.log
timeout 600 python bench.py --steps 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -5 gpurun_out/r2c_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("C2 ms/step %.1f value %.3g e2e %.3g frac %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]),
      {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()})
for k, v in d.get("extra", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"])
    else:
        print(k, "ms %.1f value %.3g" % (v["ms_per_step"], v["value"]), "e2e %.3g" % v.get("e2e", {}).get("value", 0),
              {n: round(x["ms"], 2) for n, x in v.get("roofline", {}).get("kernels", {}).items()} if "roofline" in v else v.get("kernels"),
              "coll", v.get("roofline", {}).get("collectives_ms"), "vs_cpu", v.get("vs_cpu_1core"))
PY
for v in "9 160" "7 192"; do
  set -- $v
  echo "== ws M=$1 NT=$2"
  BLG_WS_M=$1 BLG_WS_NT=$2 timeout 300 python bench.py --steps 3 --no-extra --no-cpu-baseline 2> gpurun_out/r2c_ws_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.1f' % d['ms_per_step'], {k: round(v['ms'], 2) for k, v in d['roofline']['kernels'].items()})"
done
