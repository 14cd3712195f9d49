#!/bin/bash
# cluster-stream decode (mode 4): parity tests, then timing against mode 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15 > gpurun_out/cs_tests.log
cat gpurun_out/cs_tests.log
for b in 1 8 32; do
  timeout 120 python scripts/profile_step.py --batch $b --steps 500 --mode 4 --tc 1 2>&1 | tail -1
done | tee gpurun_out/cs_perf.log
timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --mode 4 --tc 1 2>&1 | tail -1 | tee -a gpurun_out/cs_perf.log
timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --mode 1 --tc 1 2>&1 | tail -1 | tee -a gpurun_out/cs_perf.log
timeout 200 python scripts/timeline_cs.py --batch 1 > gpurun_out/tl_cs_b1.log 2>&1; tail -8 gpurun_out/tl_cs_b1.log
timeout 200 python scripts/timeline_cs.py --batch 32 > gpurun_out/tl_cs_b32.log 2>&1; tail -8 gpurun_out/tl_cs_b32.log
