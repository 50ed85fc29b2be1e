#!/bin/bash
# smoke() after the family fix + sanitizer over the new 1-D kernels
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/round2/r2K_sanitize.sh
