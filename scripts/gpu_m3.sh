#!/bin/bash
# large-batch decode on the tensor cores (mode 3): parity tests, then its step time against the chunked cluster-stream kernel (mode 5)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "tcgen05 or large_batch or cfg2_b32" 2>&1 | tail -3
for b in 64 128 256; do
  timeout 200 python scripts/profile_step.py --batch $b --steps 300 --mode 1 --tc 1 --tcmin 64 2>&1 | tail -1
  timeout 200 python scripts/profile_step.py --batch $b --steps 300 --mode 5 --tc 1 2>&1 | tail -1
done
timeout 200 python scripts/profile_step.py --batch 128 --steps 100 --mode 1 --tc 1 --tcmin 64 --prompt 600 --lo 300 --hi 300 2>&1 | tail -1
timeout 200 python scripts/profile_step.py --batch 128 --steps 100 --mode 5 --tc 1 --prompt 600 --lo 300 --hi 300 2>&1 | tail -1
