#!/bin/bash
# 8 x B200 over NCCL: the bench under torchrun (C2 weak; C3 / C4 / C5 strong under extra)
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2/r2Y2_8gpu.sh'
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 8 --steps 3 > gpurun_out/r2Y2_bench_n8.json 2> gpurun_out/r2Y2_bench_n8.err
echo "torchrun exit code $?"; tail -5 gpurun_out/r2Y2_bench_n8.err | cut -c1-300
python - <<PY
import json
d = [json.loads(l) for l in open("gpurun_out/r2Y2_bench_n8.json") if l.startswith("{")][0]
print("N=8 C2 ms/step %.1f value %.3g e2e %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]),
      "kernel share %.3f coll %.4f" % (d["roofline"]["kernel_share_of_step"], d["roofline"]["collective_share_of_step"]))
for k, v in d.get("extra", {}).items():
    if "error" in v:
        print(k, "ERROR", v["error"]); continue
    r = v.get("roofline", {})
    print(k, "ms %.1f value %.3g e2e %.3g" % (v["ms_per_step"], v["value"], v.get("e2e", {}).get("value", 0)),
          "coll_ms", r.get("collectives_ms"), "coll share", r.get("collective_share_of_step"), v.get("sharing", {}).get("executed_over_nominal"))
PY
