#!/bin/bash
# wide decode kernel (mode 6): parity against the reference goldens, then timing against the cluster-stream kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_horizons.py -q -s -x -k "6" > gpurun_out/r2_wide_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_wide_tests.log
grep -E "max \|dlogit|passed|failed|Error|error|assert" gpurun_out/r2_wide_tests.log | head -30
for b in 1 2 4 8; do for m in 6 4; do timeout 120 python scripts/profile_step.py --batch $b --steps 500 --tc 1 --mode $m 2>&1 | tail -1; done; done | tee gpurun_out/r2_wide_prof.log
