#!/bin/bash
cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so
echo "== base"; timeout 200 python scripts/timeline_cs.py --batch 1 2>&1 | grep -A2 "CTA 5" | head -3
for lib in "$@"; do
  cp $lib gpt-sovits_b200/libt2s_b200.so
  echo "== $lib"; timeout 200 python scripts/timeline_cs.py --batch 1 2>&1 | grep -A2 "CTA 5" | head -3
done
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
