#!/bin/bash
cp gpt-sovits_b200/libt2s_b200.so /tmp/base.so
cp gpt-sovits_b200/libt2s_probe.so gpt-sovits_b200/libt2s_b200.so
python scripts/attn_probe.py 2>&1 | tail -3
cp /tmp/base.so gpt-sovits_b200/libt2s_b200.so
