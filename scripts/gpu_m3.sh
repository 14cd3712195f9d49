#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05 or large_batch" 2>&1 | tail -3
for b in 128 256; do
  timeout 200 python scripts/profile_step.py --batch $b --steps 300 --mode 1 --tc 1 --tcmin 64 2>&1 | tail -1
  timeout 200 python scripts/profile_step.py --batch $b --steps 300 --mode 5 --tc 1 2>&1 | tail -1
done
