"""Config 3 on one GPU through StreamingSession: job time against slice length and admission granularity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic
import bench
c = bench.CFG3
dev = torch.device("cuda:0")
sd = synthetic.make_state_dict(seed=c["weight_seed"], eos_scale=c["eos_scale"])
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG, device=dev); eng.load_state_dict(sd, pe=synthetic.sine_pe())
n = c["total"]
L = synthetic.config_lens(n, c["lo"], c["hi"], seed=300)
ids, lens, prompt, bert = synthetic.make_inputs(n, L, c["prompt"], seed=301)
ids = [t.to(dev) for t in ids]; bert = [t.to(dev) for t in bert]; prompt = prompt.to(dev)
kw = dict(top_k=c["top_k"], top_p=c["top_p"], temperature=c["temperature"], repetition_penalty=c["repetition_penalty"],
          early_stop_num=c["cap"], eos_suppress_steps=c["eos_window"], max_steps=1500, seed=77)
def run(slots, sl, am):
    sess = gsb.StreamingSession(eng, slots=slots, positions=c["hi"] + c["prompt"] + c["cap"] + 8, slice_steps=sl, admit_min=am, **kw)
    for i in range(n):
        sess.submit([ids[i]], [bert[i]], prompt[i:i + 1])
    tot = 0
    for key, toks, k in sess:
        tot += int(k)
    return tot
for slots, sl, am in [(56, 24, 8), (56, 16, 8), (56, 32, 8), (56, 48, 8), (56, 24, 4), (56, 24, 16), (56, 32, 16), (49, 24, 8), (42, 24, 8)]:
    run(slots, sl, am)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tot = run(slots, sl, am)
    torch.cuda.synchronize(); ms = 1000 * (time.perf_counter() - t0)
    st = eng.stats()
    print(f"slots {slots} slice {sl} admit_min {am}: {ms:.1f} ms wall, {tot} tokens, {tot/ms:.1f} k tok/s; engine prefill {st['prefill_ms']:.1f} ms decode {st['decode_ms']:.1f} ms, {int(st['decode_steps'])} steps")
