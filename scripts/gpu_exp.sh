#!/bin/bash
cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so
for v in 1 2 3; do
  cp gpt-sovits_b200/libt2s_kvpf$v.so gpt-sovits_b200/libt2s_b200.so
  echo "=== KVPF $v"
  python scripts/profile_step.py --steps 1000 | tail -1
done
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
