#!/bin/bash
# compute-sanitizer over the kernels that hand-roll mbarrier / named-barrier / DSMEM protocols:
# memcheck, racecheck (shared-memory hazards inside a CTA) and synccheck (barrier misuse) on small cases of every
# kernel family -- warp-specialised 1-D (fast1d_ws), cluster-resident 2-D (cluster2d), tiled online step (online2d),
# generic resident / stream kernels.   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Output: gpurun_out/sanitize_<tool>.log + a summary line per tool (copied to profiles/ by hand).
mkdir -p gpurun_out
CASES='test_case_matches_reference_golden and (ref_coal_config1 or syn_hyper_poisson_sweep or ref_tm_nested or syn_online_mixed or ref_om_laplace or syn_hyper_dead_combo) or test_cluster2d_on_golden_cases and syn_cps_gauss_2d or test_online2d_on_golden_cases and syn_online_mixed or test_stream_kernels_on_golden_cases and syn_hyper_gauss_2d or test_cluster2d_matches_cpu_oracle and gauss_2d_100x100 and full'
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
    echo "== $tool"
    timeout 1200 $SAN --tool $tool --error-exitcode 9 --print-limit 20 --log-file gpurun_out/sanitize_$tool.raw \
        python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$CASES" > gpurun_out/sanitize_$tool.log 2>&1
    echo "exit code $?" >> gpurun_out/sanitize_$tool.log
    tail -3 gpurun_out/sanitize_$tool.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:" gpurun_out/sanitize_$tool.raw | sort | uniq -c | sort -rn | head -8
done
