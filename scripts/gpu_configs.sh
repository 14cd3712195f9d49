#!/bin/bash
# BASELINE.json configs 1, 4 (batch sweep), 5 on one GPU; correctness smoke at large batch.
mkdir -p gpurun_out
: > gpurun_out/configs.log
python scripts/profile_step.py --batch 1 --lo 80 --hi 80 --steps 500 --tc 1 --reps 2 | tail -1 >> gpurun_out/configs.log
for b in 8 64 148 149 256; do python scripts/profile_step.py --batch $b --steps 500 --tc 1 --reps 1 | tail -1 >> gpurun_out/configs.log; done
python scripts/profile_step.py --batch 128 --lo 300 --hi 300 --prompt 600 --steps 600 --tc 1 --reps 1 | tail -1 >> gpurun_out/configs.log
cat gpurun_out/configs.log
python - <<'PY'
# large-batch parity: B=160 (> number of SMs: two split-KV descriptors per CTA) teacher-forced vs B=1 runs of a few rows
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic
sd = synthetic.make_state_dict(seed=0); eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=synthetic.sine_pe())
B = 160
L = synthetic.config_lens(B, 20, 70, seed=5)
ids, lens, prompt, bert = synthetic.make_inputs(B, L, 40, seed=6)
ids = [t.cuda() for t in ids]; bert = [t.cuda() for t in bert]; prompt = prompt.cuda()
n = 12
forced = torch.randint(0, 1024, (B, n), dtype=torch.int32)
r = eng.infer(ids, bert, prompt, top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced, capture_logits=n)
big = r.logits.cpu().numpy()
worst = 0.0
for b in (0, 77, 147, 148, 159):
    r1 = eng.infer([ids[b]], [bert[b]], prompt[b:b+1], top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced[b:b+1], capture_logits=n)
    d = float(np.abs(r1.logits.cpu().numpy()[:, 0, :1024] - big[:, b, :1024]).max()); worst = max(worst, d)
    print("slot", b, "max |dlogit| batch160 vs alone:", round(d, 5))
assert worst <= 0.06, worst
print("large-batch invariance OK, idx sample:", r.idx[:4])
PY
