#!/bin/bash
# A/B: default library vs a variant library (path in $1)
for lib in base variant; do
  if [ $lib == variant ]; then cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so; cp $1 gpt-sovits_b200/libt2s_b200.so; fi
  echo "== $lib"
  for b in 1 8; do timeout 120 python scripts/profile_step.py --batch $b --steps 500 --tc 1 2>&1 | tail -1; done
  timeout 120 python scripts/profile_step.py --batch 32 --steps 1000 --tc 1 2>&1 | tail -1
done
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
