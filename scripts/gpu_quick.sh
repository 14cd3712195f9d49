#!/bin/bash
# quick regression + timing of the default decode path, plus DRAM traffic of the bench decode launch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for b in 1 8; do timeout 120 python scripts/profile_step.py --batch $b --steps 500 --tc 1 2>&1 | tail -1; done
timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1
if [ "$1" == "dram" ]; then
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_decode_cluster -c 1 \
   python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | grep -E "dram__|lts__" 
fi
