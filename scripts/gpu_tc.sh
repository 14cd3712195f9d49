#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"
grep -E "tcgen05 prefill|passed|failed|Error|error|assert " gpurun_out/pytest_tc.log | head -20
