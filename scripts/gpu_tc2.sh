#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "large_batch or modes_identical or full_size" > gpurun_out/pytest_tc2.log 2>&1; echo "pytest rc=$?"
grep -E "tcgen05 decode|passed|failed|Error|error|assert " gpurun_out/pytest_tc2.log | head -20
for b in 64 128 256; do python scripts/profile_step.py --batch $b --steps 300 --tc 1 | tail -1; done
python scripts/profile_step.py --batch 128 --lo 300 --hi 300 --prompt 600 --steps 600 --tc 1 | tail -1
