#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "max \|dlogit|agree|mismatch|invariance|passed|failed|Error|error|assert|continuous|stream" gpurun_out/pytest_gpu.log | head -80
