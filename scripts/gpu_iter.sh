#!/bin/bash
# one optimisation iteration: parity tests, then timelines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|assert " gpurun_out/pytest_gpu.log | head -20
timeout 300 python scripts/timeline.py --batch 32 --layers 1 2>&1 | tail -14
timeout 300 python scripts/timeline.py --batch 1 --lo 80 --hi 80 --layers 1 2>&1 | tail -13
timeout 300 python scripts/profile_step.py --mode 1 --steps 1000 2>&1 | tail -1
timeout 300 python scripts/profile_step.py --mode 1 --batch 1 --lo 80 --hi 80 --steps 500 2>&1 | tail -1
