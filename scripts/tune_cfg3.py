"""Pick the EOS-row scale for bench.py's config 3 (greedy, natural EOS): mean kept tokens per utterance by scale."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic
L = synthetic.config_lens(120, 60, 140, seed=300)
ids, lens, prompt, bert = synthetic.make_inputs(120, L, 150, seed=301)
ids = [t.cuda() for t in ids]; bert = [t.cuda() for t in bert]; prompt = prompt.cuda()
for seed in (3, 5):
    for scale in (1.0, 1.1, 1.2, 1.3, 1.4):
        sd = synthetic.make_state_dict(seed=seed, eos_scale=scale)
        eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=synthetic.sine_pe())
        r = eng.infer(ids, bert, prompt, top_k=1, early_stop_num=1000, eos_suppress_steps=1)
        idx = np.array(r.idx)
        print(f"weight seed {seed} eos_scale {scale}: greedy idx mean {idx.mean():.1f} min {idx.min()} max {idx.max()} median {np.median(idx):.0f} capped {(idx >= 1000).sum()}", flush=True)
        eng.close()
