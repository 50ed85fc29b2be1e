#!/bin/bash
# the other geometries of the DMMA kernels (tiles per warp 1 / 6, 8 compute warps) against the C oracle
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "poisson_grid" > gpurun_out/r2geo_pytest.log 2>&1
tail -12 gpurun_out/r2geo_pytest.log | cut -c1-250
