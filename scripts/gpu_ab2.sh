#!/bin/bash
# A/B: variants (paths in args) at the bench shape only
cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so
echo "== base"; timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1
for lib in "$@"; do
  cp $lib gpt-sovits_b200/libt2s_b200.so
  echo "== $lib"; timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1
done
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
