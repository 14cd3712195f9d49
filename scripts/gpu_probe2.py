import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
sd = synthetic.make_state_dict(seed=0); pe = synthetic.sine_pe()
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=pe)
g = np.load("tests/golden/reffree_b1.npz")
ids, lens, _, bert = synthetic.make_inputs(1, [48], 0, seed=4)
ids = [t.cuda() for t in ids]; bert = [t.cuda() for t in bert]
n = g["logits"].shape[0]
for det in (1, 0):
  for mode in (0, 1):
    eng.set_option(_lib.OPT_DECODE_MODE, mode)
    r = eng.infer(ids, bert, None, top_k=1, early_stop_num=20, eos_suppress_steps=11, capture_logits=n)
    got = r.logits.cpu().numpy()
    d = [float(np.abs(got[s,0,:1024]-g["logits"][s,0,:1024]).max()) for s in range(n)]
    print("det", det, "mode", mode, "idx", r.idx, "per-step:", np.round(d, 3))
    print("   tokens", r.sequences()[0].cpu().numpy()[:21], "ref", g["y"][0][:21])
