#!/bin/bash
# compute-sanitizer over the FINAL DMMA kernels (lean loops): memcheck + racecheck on small cases
mkdir -p gpurun_out
CASES='test_case_matches_reference_golden and (ref_coal_config1 or syn_hyper_poisson_sweep or syn_hyper_dead_combo or syn_hyper_poisson_forward_only) or test_cuda_matches_cpu_oracle and (poisson_wide_kernels or poisson_grid_250)'
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
    echo "== $tool"
    timeout 400 $SAN --tool $tool --error-exitcode 9 --print-limit 20 --log-file gpurun_out/r2san2_$tool.raw \
        python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$CASES" > gpurun_out/r2san2_$tool.log 2>&1
    echo "exit code $?" >> gpurun_out/r2san2_$tool.log
    tail -3 gpurun_out/r2san2_$tool.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:" gpurun_out/r2san2_$tool.raw | sort | uniq -c | sort -rn | head -8
done
