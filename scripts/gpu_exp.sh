#!/bin/bash
for v in T2S_EXP_VEC_LATE; do
  cp gpt-sovits_b200/libt2s_$v.so gpt-sovits_b200/libt2s_b200.so
  echo "=== $v"
  python scripts/timeline.py --batch 1 --lo 80 --hi 80 --layers 1 2>&1 | grep -E "mean phase|step total|qkv L1" -A1 | grep -v "^--"
  python scripts/timeline.py --batch 32 --layers 1 2>&1 | grep -E "mean phase|step total|qkv L1" -A1 | grep -v "^--"
done
