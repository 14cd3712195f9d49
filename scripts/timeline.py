"""Per-phase timeline of one persistent decode step: work / wait cycles per CTA per barrier."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32); ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--at", type=int, default=40); ap.add_argument("--prompt", type=int, default=150)
ap.add_argument("--lo", type=int, default=60); ap.add_argument("--hi", type=int, default=120)
ap.add_argument("--layers", type=int, default=3, help="layers to print in detail")
a = ap.parse_args()
sd = synthetic.make_state_dict(seed=0, eos_scale=0.0)
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=synthetic.sine_pe())
nb = 24 * 5 + 3
ncta = int(eng.stats()["num_sms"])
tl = torch.zeros((ncta, nb, 2), dtype=torch.int64, device="cuda")
_lib.check(eng.lib.t2s_set_timeline(eng._h, tl.data_ptr(), a.at, nb))
L = synthetic.config_lens(a.batch, a.lo, a.hi, seed=100)
ids, lens, prompt, bert = synthetic.make_inputs(a.batch, L, a.prompt, seed=200)
r = eng.infer([t.cuda() for t in ids], [t.cuda() for t in bert], prompt.cuda(), top_k=15, early_stop_num=a.steps, seed=1)
st = r.stats
print(f"B={a.batch}: {1000*st['decode_ms']/st['decode_steps']:.1f} us/step over {int(st['decode_steps'])} steps")
t = tl.cpu().numpy().astype(np.float64)
arrive, release = t[:, :, 0], t[:, :, 1]
work = np.empty_like(arrive); work[:, 1:] = arrive[:, 1:] - release[:, :-1]; work[:, 0] = np.nan
wait = release - arrive
names = ["qkv", "attn", "oproj", "ffn1", "ffn2"]
MHZ = 1965.0
def us(c): return c / MHZ
print("phase       work_max  work_med  busyCTAs  wait_min(barrier)  phase_total(us)")
tot = {}
for k in range(nb):
    nm = names[k % 5] if k < 120 else ["head", "sample", "plan"][k - 120]
    w = work[:, k]
    total = (release[:, k] - (release[:, k-1] if k > 0 else arrive[:, k])).mean()
    tot.setdefault(nm, []).append(us(total))
    if k < 5 * a.layers or k >= 120:
        print(f"{k:3d} {nm:7s} {us(np.nanmax(w)) if k else 0:8.2f} {us(np.nanmedian(w)) if k else 0:8.2f} {int((w > 0.3 * np.nanmax(w)).sum()) if k else 0:8d} "
              f"{us(wait[:, k].min()):10.2f} {us(total):14.2f}")
print("mean phase time (us):", {k: round(float(np.mean(v)), 2) for k, v in tot.items()})
print("step total (us):", round(sum(sum(v) for v in tot.values()), 1))

