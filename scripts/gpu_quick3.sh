#!/bin/bash
# regression (all GPU tests) + prefill / decode timing with the persistent (tc 1) and one-tile-per-CTA (tc 2) prefill GEMMs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for tc in 1 2; do
  timeout 120 python scripts/profile_step.py --batch 32 --steps 200 --tc $tc 2>&1 | tail -1
  timeout 120 python scripts/profile_step.py --batch 56 --steps 50 --tc $tc --prompt 600 --lo 300 --hi 300 2>&1 | tail -1
done
