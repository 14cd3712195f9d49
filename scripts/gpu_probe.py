"""Determinism / batch-invariance probe (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
sd = synthetic.make_state_dict(seed=0); pe = synthetic.sine_pe()
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=pe)
L = [37, 64, 65, 20, 128]
ids, lens, prompt, bert = synthetic.make_inputs(5, L, 77, seed=11)
ids = [t.cuda() for t in ids]; bert = [t.cuda() for t in bert]; prompt = prompt.cuda()
n = 10
forced = torch.randint(0, 1024, (5, n), dtype=torch.int32)
def run(sel, mode):
    eng.set_option(_lib.OPT_DECODE_MODE, mode)
    r = eng.infer([ids[i] for i in sel], [bert[i] for i in sel], prompt[sel], top_k=1, early_stop_num=n-1,
                  eos_suppress_steps=1, forced=forced[sel], capture_logits=n)
    return r.logits.cpu().numpy()
for mode in (0, 1):
    a = run([0,1,2,3,4], mode); b = run([0,1,2,3,4], mode)
    print("mode", mode, "same call twice, per-step max diff:", np.round(np.abs(a-b)[:, :, :1024].max(axis=(1,2)), 6))
    c = run([2], mode)
    print("mode", mode, "batch vs single (utt 2) per-step:", np.round(np.abs(a[:, 2, :1024]-c[:, 0, :1024]).max(axis=1), 6))
    d = run([2], mode)
    print("mode", mode, "single twice:", np.round(np.abs(d[:, 0, :1024]-c[:, 0, :1024]).max(axis=1), 6))
    e2 = run([1,2], mode)
    print("mode", mode, "pair vs single (utt 2) per-step:", np.round(np.abs(e2[:, 1, :1024]-c[:, 0, :1024]).max(axis=1), 6))
