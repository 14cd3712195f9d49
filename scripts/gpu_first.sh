#!/bin/bash
# First GPU contact: smoke, parity tests, short bench.  Everything under `timeout`.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 --cpu-steps 8 > gpurun_out/bench_first.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/bench_first.log
tail -5 gpurun_out/bench_first.log
