#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_step.py --barrier-bench --steps 200 --reps 2 > gpurun_out/prof_plain.log 2>&1
for m in 0 2; do python scripts/profile_step.py --mode $m --steps 200 --reps 2 >> gpurun_out/prof_plain.log 2>&1; done
python scripts/profile_step.py --mode 1 --det 0 --steps 200 --reps 2 >> gpurun_out/prof_plain.log 2>&1
python scripts/profile_step.py --mode 1 --batch 1 --lo 80 --hi 80 --steps 300 --reps 2 >> gpurun_out/prof_plain.log 2>&1
python scripts/profile_step.py --mode 1 --batch 8 --steps 300 --reps 2 >> gpurun_out/prof_plain.log 2>&1
python scripts/profile_step.py --mode 1 --batch 128 --steps 100 --reps 2 >> gpurun_out/prof_plain.log 2>&1
grep -E "barrier bench|mode" gpurun_out/prof_plain.log
# per-launch device times of two full decode steps (graph-less mode 2) at step ~100
python scripts/profile_step.py --mode 2 --steps 110 > gpurun_out/ncu_plain_check.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_phase -s 12299 -c 244 --csv --log-file gpurun_out/launches_decode_b32.csv python scripts/profile_step.py --mode 2 --steps 110 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
# prefill launch list
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"k_phase|k_prefill|k_bert|k_embed" -c 140 --csv --log-file gpurun_out/launches_prefill_b32.csv python scripts/profile_step.py --mode 2 --steps 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
