"""Decode-step profiling driver: B utterances, prefill, then N decode steps in the chosen mode.
Used under ncu for the per-launch list (mode 2) and the --set full capture (mode 1)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpt_sovits_b200 as gsb
from gpt_sovits_b200 import synthetic, _lib
import ctypes as C
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--mode", type=int, default=5)
ap.add_argument("--prompt", type=int, default=150)
ap.add_argument("--lo", type=int, default=60)
ap.add_argument("--hi", type=int, default=120)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--det", type=int, default=1)
ap.add_argument("--tc", type=int, default=0)
ap.add_argument("--barrier-bench", action="store_true")
ap.add_argument("--tcmin", type=int, default=-1, help="T2S_OPT_TC_DECODE_MIN_BATCH (-1: library default)")
a = ap.parse_args()
sd = synthetic.make_state_dict(seed=0, eos_scale=0.0)
eng = gsb.T2SEngine(synthetic.S1V2_CONFIG); eng.load_state_dict(sd, pe=synthetic.sine_pe())
eng.set_option(_lib.OPT_DECODE_MODE, a.mode); eng.set_option(_lib.OPT_PREFILL_GEMM, a.tc);
if a.tcmin >= 0: eng.set_option(_lib.OPT_TC_DECODE_MIN_BATCH, a.tcmin)
if a.barrier_bench:
    for ncta in (148, 74, 37):
        for n in (1000, 10000):
            ms = C.c_float(0)
            _lib.check(eng.lib.t2s_bench_barrier(eng._h, n, ncta, C.byref(ms), None))
            print(f"barrier bench: {ncta} CTAs, {n} barriers: {ms.value*1000/n:.3f} us/barrier")
L = synthetic.config_lens(a.batch, a.lo, a.hi, seed=100)
ids, lens, prompt, bert = synthetic.make_inputs(a.batch, L, a.prompt, seed=200)
ids = [t.cuda() for t in ids]; bert = [t.cuda() for t in bert]; prompt = prompt.cuda()
for rep in range(a.reps):
    r = eng.infer(ids, bert, prompt, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                  early_stop_num=a.steps, eos_suppress_steps=1, seed=1 + rep)
    st = r.stats
    by = st["decode_steps"] * st["weight_bytes_per_step"] + st["kv_bytes_per_position"] * (st["decode_kv_positions"] + st["decode_tokens"])
    print(f"mode {a.mode} tc {a.tc} B={a.batch}: prefill {st['prefill_ms']:.2f} ms ({int(st['prefill_rows'])} rows), decode {st['decode_ms']:.2f} ms / "
          f"{int(st['decode_steps'])} steps = {1000*st['decode_ms']/max(st['decode_steps'],1):.1f} us/step, "
          f"{by/st['decode_ms']/1e6:.0f} GB/s algorithmic, mean KV {st['decode_kv_positions']/max(st['decode_tokens'],1):.0f}")
