#!/bin/bash
# round 2, first GPU call: the whole GPU suite (new horizon goldens included) + timing of the default path at batch 1 / 8 / 32
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_first.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_first.log
grep -E "max \|dlogit|agree|mismatch|passed|failed|Error|error|assert|skipped" gpurun_out/r2_pytest_first.log | head -80
for b in 1 8; do timeout 120 python scripts/profile_step.py --batch $b --steps 500 --tc 1 2>&1 | tail -1; done | tee gpurun_out/r2_prof_first.log
timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1 | tee -a gpurun_out/r2_prof_first.log
