"""Marker timeline of one wide decode step (mode 6): mean time between consecutive markers of a layer, per CTA role."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1); ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--at", type=int, default=40); ap.add_argument("--prompt", type=int, default=150)
ap.add_argument("--lo", type=int, default=60); ap.add_argument("--hi", type=int, default=120)
a = ap.parse_args()
sd = synthetic.make_state_dict(seed=0, eos_scale=0.0)
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=synthetic.sine_pe())
eng.set_option(_lib.OPT_DECODE_MODE, 6)
PER = 9  # markers per layer
nm = 2 + 24 * PER + 1 + 5
slots = (nm + 1) // 2
ncta = 144
tl = torch.zeros((ncta + 1, slots * 2), dtype=torch.int64, device="cuda")
_lib.check(eng.lib.t2s_set_timeline(eng._h, tl.data_ptr(), a.at, slots))
L = synthetic.config_lens(a.batch, a.lo, a.hi, seed=100)
ids, lens, prompt, bert = synthetic.make_inputs(a.batch, L, a.prompt, seed=200)
r = eng.infer([t.cuda() for t in ids], [t.cuda() for t in bert], prompt.cuda(), top_k=15, early_stop_num=a.steps, seed=1)
st = r.stats
print(f"B={a.batch}: {1000*st['decode_ms']/st['decode_steps']:.1f} us/step over {int(st['decode_steps'])} steps, mode {int(st['decode_mode'])}")
t = tl.cpu().numpy().astype(np.float64)
MHZ = 1965.0
names = ["x4 gather+ln2", "q gemv+epi", "attention", "publish/merge", "wo", "x2 gather+ln1", "w1+publish", "x3 gather", "w2+publish"]
for cta, role in ((0, "s=0 (k)"), (1, "s=1 (v)"), (3, "s=3 (w2)"), (6, "s=6 (wo)"), (8, "s=8 (merger)"), (75, "h=8 s=3")):
    row = t[cta]
    if row[1] == 0: continue
    d = np.diff(row[:nm]) / MHZ
    print(f"CTA {cta} {role}: prologue {d[0]:.2f} us")
    lay = d[1: 1 + 24 * PER].reshape(24, PER)
    print("  per-layer mean (us):", {n: round(float(v), 2) for n, v in zip(names, lay.mean(axis=0))})
    print("  layer total mean %.2f us, x24 = %.1f us; layer 0: %s; layer 5: %s" % (lay.sum(axis=1).mean(), lay.sum(), np.round(lay[0], 2), np.round(lay[5], 2)))
    tail = d[1 + 24 * PER:]
    print("  tail (x4 gather+ln (head unit), head gemv, gbar1, sample, gbar2):", np.round(tail, 2))
