#!/bin/bash
# 2-GPU bench with diagnostics (the first attempt of the session ended without output)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus 2 --steps 3 > gpurun_out/r2J_bench_n2.json 2> gpurun_out/r2J_bench_n2.err
echo "torchrun exit code $?"
tail -25 gpurun_out/r2J_bench_n2.err | cut -c1-300
wc -c gpurun_out/r2J_bench_n2.json
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2J_bench_n2.json") if l.startswith("{")][0]
print("N=2 C2 ms/step %.1f value %.3g e2e %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]),
      "kernel share %.3f coll %.4f" % (d["roofline"]["kernel_share_of_step"], d["roofline"]["collective_share_of_step"]))
for k, v in d.get("extra", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"]); continue
    r = v.get("roofline", {})
    print(k, "ms %.1f value %.3g e2e %.3g" % (v["ms_per_step"], v["value"], v.get("e2e", {}).get("value", 0)),
          "coll_ms", r.get("collectives_ms"), "coll share", r.get("collective_share_of_step"), v.get("sharing", {}).get("executed_over_nominal"))
PY
