#!/bin/bash
cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so
for v in base acq; do
  [ $v = acq ] && cp gpt-sovits_b200/libt2s_acq.so gpt-sovits_b200/libt2s_b200.so
  echo "=== $v"
  python scripts/profile_step.py --barrier-bench --steps 1000 2>&1 | grep -E "148 CTAs, 10000|mode 1"
  python scripts/profile_step.py --batch 1 --lo 80 --hi 80 --steps 500 | tail -1
done
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
