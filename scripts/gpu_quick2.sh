#!/bin/bash
# regression (all GPU tests) + timing of the default decode path at batch 1 / 8 / 32
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for b in 1 8; do timeout 120 python scripts/profile_step.py --batch $b --steps 500 --tc 1 2>&1 | tail -1; done
timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1
timeout 120 python scripts/profile_step.py --batch 56 --steps 600 --tc 1 --prompt 600 --lo 300 --hi 300 2>&1 | tail -1
