#!/bin/bash
# compute-sanitizer over the 1-D kernels of this session: DMMA kernels (fast1d_mma, incl. a dead combo and pacing) and the
# interleaved variant (fast1d_il: per-chain counters + fences instead of CTA barriers)
mkdir -p gpurun_out
CASES='test_case_matches_reference_golden and (ref_coal_config1 or syn_hyper_poisson_sweep or syn_hyper_dead_combo or syn_hyper_poisson_forward_only) or test_cuda_matches_cpu_oracle and (poisson_interleaved_small_grid or poisson_wide_kernels)'
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
    echo "== $tool"
    timeout 900 $SAN --tool $tool --error-exitcode 9 --print-limit 20 --log-file gpurun_out/r2K_sanitize_$tool.raw \
        python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$CASES" > gpurun_out/r2K_sanitize_$tool.log 2>&1
    echo "exit code $?" >> gpurun_out/r2K_sanitize_$tool.log
    tail -3 gpurun_out/r2K_sanitize_$tool.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:" gpurun_out/r2K_sanitize_$tool.raw | sort | uniq -c | sort -rn | head -8
done
