import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
g = np.load("/root/repo/tests/golden/retire_b6.npz", allow_pickle=True)
sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), eos_scale=float(g["eos_scale"]))
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG, device="cuda:0"); eng.load_state_dict(sd, pe=synthetic.sine_pe())
sys.path.insert(0, "/root/repo/tests")
from test_gpu_parity import _inputs
ids, bert, prompt = _inputs(g)
P = int(g["prompt_len"])
forced = torch.from_numpy(g["y"][:, P:P + 8].copy()).to(torch.int32).clamp(min=0)
out = {}
for gm in (1, 2, 0):
    eng.set_option(_lib.OPT_PREFILL_GEMM, gm)
    r = eng.infer(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=1, forced=forced, capture_logits=8)
    out[gm] = r.logits.cpu().numpy()
for a, b in ((1, 2), (1, 0), (2, 0)):
    d = np.nanmax(np.abs(out[a][:, :, :1024] - out[b][:, :, :1024]), axis=(1, 2))
    print(f"prefill gemm {a} vs {b}: max |dlogit| per step", np.round(d, 4))
