#!/bin/bash
# Closing GPU run of round 1 (one gpurun call): full parity suite, the opt-in tiled online kernels (tests + A/B on the
# C5 shape), smoke, bench, one ncu capture of the tile kernel.  Everything lands in gpurun_out/r1k_*.
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/r1k_pytest_gpu.log 2>&1
tail -4 gpurun_out/r1k_pytest_gpu.log
BLG_TEST_ONLINE2D=1 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k online2d > gpurun_out/r1k_pytest_online2d.log 2>&1
tail -12 gpurun_out/r1k_pytest_online2d.log
timeout 60 python tools/exp_online.py 512 40 > gpurun_out/r1k_online_stream.log 2>&1
tail -2 gpurun_out/r1k_online_stream.log
BLG_ONLINE2D=1 timeout 60 python tools/exp_online.py 512 40 > gpurun_out/r1k_online_tiled.log 2>&1
tail -2 gpurun_out/r1k_online_tiled.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r1k_bench.json"))
print("bench ms/step %.1f value %.3g e2e %.3g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]),
      {k: round(v["ms"], 2) for k, v in d["roofline"]["kernels"].items()}, "c3", d["extra"]["c3_sample"]["ms_per_step"])
PY
BLG_ONLINE2D=1 timeout 90 ncu --set full --clock-control none --import-source on -k regex:online2d_tile -s 20 -c 1 \
    -f -o gpurun_out/r1k_online2d_tile python tools/exp_online.py 512 30 > gpurun_out/r1k_ncu_online2d.log 2>&1
tail -2 gpurun_out/r1k_ncu_online2d.log
