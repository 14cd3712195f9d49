"""Config-3 utterances through the chunked call and through the continuous-batching session: per-utterance comparison
(greedy decoding: a difference is either a near-tie flipped by a different fp32 summation order, or a bug)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic
import bench
c = bench.CFG3
dev = torch.device("cuda:0")
sd = synthetic.make_state_dict(seed=c["weight_seed"], eos_scale=c["eos_scale"])
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG, device=dev); eng.load_state_dict(sd, pe=synthetic.sine_pe())
n = int(sys.argv[1]) if len(sys.argv) > 1 else c["total"]
L = synthetic.config_lens(c["total"], c["lo"], c["hi"], seed=300)[:n]
ids, lens, prompt, bert = synthetic.make_inputs(c["total"], synthetic.config_lens(c["total"], c["lo"], c["hi"], seed=300), c["prompt"], seed=301)
ids = [t.to(dev) for t in ids[:n]]; bert = [t.to(dev) for t in bert[:n]]; prompt = prompt[:n].to(dev)
kw = dict(top_k=c["top_k"], top_p=c["top_p"], temperature=c["temperature"], repetition_penalty=c["repetition_penalty"],
          early_stop_num=c["cap"], eos_suppress_steps=c["eos_window"], max_steps=1500, seed=77)
r = eng.infer(ids, bert, prompt, utt_ids=list(range(n)), **kw)
ch = [s.cpu().numpy() for s in r.sequences()]
sess = gsb.StreamingSession(eng, slots=min(c["slots"], n), positions=c["hi"] + c["prompt"] + c["cap"] + 8, slice_steps=c["slice_steps"],
                            admit_min=c["admit_min"], **kw)
for i in range(n):
    sess.submit([ids[i]], [bert[i]], prompt[i:i + 1])
cb = [None] * n
for key, toks, k in sess:
    cb[key] = toks.cpu().numpy()
solo = {}
same = 0
for i in range(n):
    a, b = ch[i], cb[i]
    if len(a) == len(b) and np.array_equal(a, b):
        same += 1
        continue
    m = min(len(a), len(b))
    d = int(np.argmax(a[:m] != b[:m])) if not np.array_equal(a[:m], b[:m]) else m
    r1 = eng.infer([ids[i]], [bert[i]], prompt[i:i + 1], utt_ids=[i], **kw)
    s = r1.sequences()[0].cpu().numpy()
    ms = min(len(s), m)
    ds_a = int(np.argmax(s[:ms] != a[:ms])) if not np.array_equal(s[:ms], a[:ms]) else ms
    ds_b = int(np.argmax(s[:ms] != b[:ms])) if not np.array_equal(s[:ms], b[:ms]) else ms
    print(f"utt {i}: chunked len {len(a)-c['prompt']}, continuous len {len(b)-c['prompt']}, first difference at generated token {d-c['prompt']}; "
          f"solo run len {len(s)-c['prompt']} agrees with chunked up to {ds_a-c['prompt']}, with continuous up to {ds_b-c['prompt']}")
print(f"{same}/{n} utterances identical")
