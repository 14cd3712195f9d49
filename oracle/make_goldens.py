"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, fp32) on the deterministic synthetic weights of gpt-sovits_b200/synthetic.py.

Run in the build container only:   python -m oracle.make_goldens
The GPU box has no /root/reference; it checks the CUDA path and the numpy oracle against these files.

Cases (weights are regenerated from the recorded seed, not stored):
  naive_b1      infer_panel (= infer_panel_naive, t2s_model.py:814), B=1, 80 phonemes + 150 prompt,
                greedy top_k=1, rp 1.35, early_stop_num=40: per-step logits, y, idx
  batch_b4      infer_panel_batch_infer (:583), ragged B=4 (60/80/72/80 phonemes), P=150, greedy,
                early_stop_num=24: per-step logits [n_active,1025], y_list, idx_list
  retire_b6     infer_panel_batch_infer with an EOS-prone head (eos_scale=1.4): sequences retire at
                different steps; y_list, idx_list, per-step logits
  reffree_b1    infer_panel_naive with prompts=None (:849-856): y, idx(=0), logits
  sampler_kat   logits_to_probs (utils.py:147) on hand-built rows: penalty sign flip, prompt
                duplicates, top-k ties, top-p boundary, temperature clamp
Round 2 (the real horizons of BASELINE.json's configs; `python -m oracle.make_goldens long_b2 ...` makes only those):
  long_b2       infer_panel_batch_infer, B=2, 300 phonemes + 600 prompt tokens (S0 = 900: 8 K/V pages, the config-5 shape),
                greedy, 24 steps
  long_b1       infer_panel_naive, 120 phonemes + 1350 prompt tokens (S0 = 1470), 40 steps: the KV length crosses 1500
                (12 pages, partial last page)
  naive_batched_b4 / naive_batched_reffree
                infer_panel_naive_batched (:781-812), 4 ragged items on the EOS-prone head (11-step EOS window per item,
                items stop at different idx); with prompts and with prompts=None
  cfg2_b32      infer_panel_batch_infer at BASELINE config 2: B=32, 60..120 phonemes + 150 prompt, top_k=15 rp 1.35
                SAMPLED by the reference (torch seed 0), 16 steps: per-step logits [32,1025] + the tokens it emitted
  fp16w_b1      infer_panel_naive on fp16-representable (NOT bf16-representable) weights, fp32 compute: what a real s1
                checkpoint holds; bounds the error of the engine's bf16 rounding of such weights
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gpt_sovits_b200  # noqa: E402  (alias loader)
from gpt_sovits_b200 import synthetic  # noqa: E402
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _run(model, fn_name, args, kwargs):
    hook = ref_harness.SampleHook().install()
    try:
        with ref_harness.quiet(), torch.no_grad():
            y, idx = getattr(model, fn_name)(*args, **kwargs)
    finally:
        hook.remove()
    return y, idx, hook


def _pad_logits(lst):
    """Steps have 1024 or 1025 columns and a shrinking number of rows -> pad with NaN."""
    n = max(t.shape[0] for t in lst)
    out = np.full((len(lst), n, 1025), np.nan, np.float32)
    for s, t in enumerate(lst):
        out[s, : t.shape[0], : t.shape[1]] = t.numpy()
    return out


def case_naive_b1(model):
    ids, lens, prompt, bert = synthetic.make_inputs(1, [80], 150, seed=1)
    y, idx, hook = _run(model, "infer_panel",
                        (ids[0][None], lens, prompt, bert[0][None]),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=40, repetition_penalty=1.35))
    # NOTE: hook.logits are post-penalty because the reference penalises in place (utils.py:167) and
    # the hook clones BEFORE calling the original sample -> they are the raw logits.
    np.savez_compressed(os.path.join(GOLD, "naive_b1.npz"), weight_seed=0, input_seed=1,
                        phoneme_lens=[80], prompt_len=150, top_k=1, top_p=1.0, temperature=1.0,
                        repetition_penalty=1.35, early_stop_num=40,
                        logits=_pad_logits(hook.logits), y=y.numpy(), idx=idx)
    print("naive_b1: idx", idx, "y", tuple(y.shape))


def case_batch_b4(model):
    L = [60, 80, 72, 80]
    ids, lens, prompt, bert = synthetic.make_inputs(4, L, 150, seed=2)
    y, idx, hook = _run(model, "infer_panel_batch_infer", (ids, lens, prompt, bert),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=24,
                             repetition_penalty=1.35, max_len=80))
    np.savez_compressed(os.path.join(GOLD, "batch_b4.npz"), weight_seed=0, input_seed=2,
                        phoneme_lens=L, prompt_len=150, top_k=1, top_p=1.0, temperature=1.0,
                        repetition_penalty=1.35, early_stop_num=24,
                        logits=_pad_logits(hook.logits),
                        y=np.stack([t.numpy() for t in y]), idx=np.array(idx))
    print("batch_b4: idx", idx)


def case_retire_b6():
    sd = synthetic.make_state_dict(seed=3, eos_scale=1.4)
    model = ref_harness.build_reference_model(sd, synthetic.S1V2_CONFIG)
    L = [40, 64, 52, 33, 64, 47]
    ids, lens, prompt, bert = synthetic.make_inputs(6, L, 60, seed=3)
    y, idx, hook = _run(model, "infer_panel_batch_infer", (ids, lens, prompt, bert),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=30,
                             repetition_penalty=1.35, max_len=64))
    ymax = max(t.shape[0] for t in y)
    ypad = np.full((6, ymax), -1, np.int64)
    for i, t in enumerate(y):
        ypad[i, : t.shape[0]] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, "retire_b6.npz"), weight_seed=3, eos_scale=1.4, input_seed=3,
                        phoneme_lens=L, prompt_len=60, top_k=1, top_p=1.0, temperature=1.0,
                        repetition_penalty=1.35, early_stop_num=30,
                        logits=_pad_logits(hook.logits), y=ypad, idx=np.array(idx))
    print("retire_b6: idx", idx)


def case_reffree_b1(model):
    ids, lens, _, bert = synthetic.make_inputs(1, [48], 0, seed=4)
    y, idx, hook = _run(model, "infer_panel_naive", (ids[0][None], lens, None, bert[0][None]),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=20, repetition_penalty=1.35))
    np.savez_compressed(os.path.join(GOLD, "reffree_b1.npz"), weight_seed=0, input_seed=4,
                        phoneme_lens=[48], prompt_len=0, top_k=1, top_p=1.0, temperature=1.0,
                        repetition_penalty=1.35, early_stop_num=20,
                        logits=_pad_logits(hook.logits), y=y.numpy().astype(np.int64), idx=idx)
    print("reffree_b1: idx", idx, "y", tuple(y.shape))


def case_sampler_kat():
    ref = ref_harness.import_reference()
    from AR.models.utils import logits_to_probs  # reference function

    rs = np.random.RandomState(7)
    rows, prevs, params, probs_out, pen_out = [], [], [], [], []

    def add(logits, prev, **kw):
        lg = torch.tensor(logits, dtype=torch.float32)[None].clone()
        pv = torch.tensor(prev, dtype=torch.int64)[None]
        p = logits_to_probs(lg, pv, **kw)
        rows.append(np.asarray(logits, np.float32))
        prevs.append(np.asarray(prev, np.int64))
        params.append([kw.get("temperature", 1.0), kw.get("top_k") or 0,
                       kw.get("top_p") if kw.get("top_p") is not None else 100.0,
                       kw.get("repetition_penalty", 1.0)])
        probs_out.append(p[0].numpy())
        pen_out.append(lg[0].numpy())  # penalised in place

    base = rs.standard_normal(1025).astype(np.float32)
    add(base, [3, 3, 7, 1000, 7], temperature=1.0, top_k=15, top_p=1.0, repetition_penalty=1.35)
    add(base, [3, 5], temperature=0.7, top_k=5, top_p=0.8, repetition_penalty=1.35)
    add(base[:1024], [0, 1023], temperature=1.0, top_k=1, top_p=1.0, repetition_penalty=1.35)
    add(base, [], temperature=0.0, top_k=20, top_p=1.0, repetition_penalty=1.0)  # clamp to 1e-5
    ties = base.copy()
    ties[[10, 20, 30, 40]] = 5.0  # 4-way tie at the top, top_k=2 keeps all four (ties kept)
    add(ties, [99], temperature=1.0, top_k=2, top_p=1.0, repetition_penalty=1.2)
    # top-p boundary: softmax = [0.64, 0.24, 0.09, 0.03]-like; cum > top_p removed except the first
    small = np.log(np.array([0.64, 0.24, 0.09, 0.03], np.float32))
    add(small, [], temperature=1.0, top_k=4, top_p=0.7, repetition_penalty=1.0)
    add(small, [], temperature=1.0, top_k=4, top_p=0.9, repetition_penalty=1.0)
    add(small, [1], temperature=1.0, top_k=3, top_p=0.0, repetition_penalty=2.0)
    neg = -np.abs(base)
    add(neg, list(range(0, 1025, 3)), temperature=1.3, top_k=100, top_p=0.95, repetition_penalty=1.5)
    add(base * 4, [1, 2, 3], temperature=1.0, top_k=2000, top_p=0.5, repetition_penalty=1.35)

    n = len(rows)
    width = np.array([r.shape[0] for r in rows])
    lg = np.full((n, 1025), np.nan, np.float32)
    pr = np.full((n, 1025), np.nan, np.float32)
    pn = np.full((n, 1025), np.nan, np.float32)
    pv = np.full((n, 400), -1, np.int64)
    for i in range(n):
        lg[i, : width[i]] = rows[i]
        pr[i, : width[i]] = probs_out[i]
        pn[i, : width[i]] = pen_out[i]
        pv[i, : len(prevs[i])] = prevs[i]
    np.savez_compressed(os.path.join(GOLD, "sampler_kat.npz"), logits=lg, width=width, prev=pv,
                        params=np.array(params, np.float64), probs=pr, penalised=pn)
    print("sampler_kat:", n, "rows")


def case_latent():
    """quantizer.decode + nearest x2 of SynthesizerTrn.decode (module/models.py:989-991), from the reference's own
    ResidualVectorQuantizer (module/quantize.py) on a small random codebook; lengths straddle the 32-wide kernel tiles."""
    import torch.nn.functional as F

    if ref_harness.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REFERENCE_ROOT)
    from module.quantize import ResidualVectorQuantizer

    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, dim, bins, T in (("a", 48, 40, 7), ("b", 80, 64, 33), ("c", 32, 16, 1)):
        q = ResidualVectorQuantizer(dimension=dim, n_q=1, bins=bins)
        cb = torch.randn(bins, dim, generator=g)
        q.vq.layers[0]._codebook.embed.copy_(cb)
        codes = torch.randint(0, bins, (1, 1, T), generator=g)
        with torch.no_grad():
            z = q.decode(codes)
            z = F.interpolate(z, size=int(z.shape[-1] * 2), mode="nearest")
        out["codebook_" + tag] = cb.numpy()
        out["codes_" + tag] = codes.numpy()
        out["latent_" + tag] = z.numpy()
    np.savez_compressed(os.path.join(GOLD, "latent.npz"), **out)
    print("latent:", {k: v.shape for k, v in out.items() if k.startswith("latent")})


def case_long_b2(model):
    L = [300, 300]
    ids, lens, prompt, bert = synthetic.make_inputs(2, L, 600, seed=12)
    y, idx, hook = _run(model, "infer_panel_batch_infer", (ids, lens, prompt, bert),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=24, repetition_penalty=1.35, max_len=300))
    np.savez_compressed(os.path.join(GOLD, "long_b2.npz"), weight_seed=0, input_seed=12, phoneme_lens=L, prompt_len=600,
                        top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35, early_stop_num=24,
                        logits=_pad_logits(hook.logits), y=np.stack([t.numpy() for t in y]), idx=np.array(idx))
    print("long_b2: idx", idx)


def case_long_b1(model):
    ids, lens, prompt, bert = synthetic.make_inputs(1, [120], 1350, seed=13)
    y, idx, hook = _run(model, "infer_panel_naive", (ids[0][None], lens, prompt, bert[0][None]),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=40, repetition_penalty=1.35))
    np.savez_compressed(os.path.join(GOLD, "long_b1.npz"), weight_seed=0, input_seed=13, phoneme_lens=[120],
                        prompt_len=1350, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35, early_stop_num=40,
                        logits=_pad_logits(hook.logits), y=y.numpy(), idx=idx)
    print("long_b1: idx", idx, "y", tuple(y.shape))


def _naive_batched(model, name, L, P, seed, stop):
    ids, lens, prompt, bert = synthetic.make_inputs(len(L), L, P, seed=seed)
    y, idx, hook = _run(model, "infer_panel_naive_batched", (ids, lens, prompt, bert),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=stop, repetition_penalty=1.35))
    # the reference loops the items: hook.logits = item 0's steps, then item 1's, ... (one row each)
    n_item = [int(t.shape[0]) - P + 1 for t in y]  # kept tokens + the stopping step
    assert sum(n_item) == len(hook.logits), (n_item, len(hook.logits))
    smax = max(n_item)
    lg = np.full((smax, len(L), 1025), np.nan, np.float32)
    sampled = np.full((len(L), smax), -1, np.int64)
    k = 0
    for b, n in enumerate(n_item):
        for s in range(n):
            t = hook.logits[k]
            lg[s, b, : t.shape[1]] = t[0].numpy()
            sampled[b, s] = int(hook.samples[k][0, 0])
            k += 1
    ymax = max(t.shape[0] for t in y)
    ypad = np.full((len(L), ymax), -1, np.int64)
    for i, t in enumerate(y):
        ypad[i, : t.shape[0]] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), weight_seed=3, eos_scale=1.4, input_seed=seed, phoneme_lens=L,
                        prompt_len=P, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35, early_stop_num=stop,
                        logits=lg, sampled=sampled, n_steps=np.array(n_item), y=ypad, idx=np.array(idx))
    print(name + ": idx", idx, "steps", n_item)


def case_naive_batched():
    sd = synthetic.make_state_dict(seed=3, eos_scale=1.4)
    model = ref_harness.build_reference_model(sd, synthetic.S1V2_CONFIG)
    _naive_batched(model, "naive_batched_b4", [40, 64, 52, 33], 60, 14, 36)
    _naive_batched(model, "naive_batched_reffree", [48, 30, 41], 0, 15, 20)


def case_cfg2_b32(model):
    L = synthetic.config_lens(32, 60, 120, seed=100)
    ids, lens, prompt, bert = synthetic.make_inputs(32, L, 150, seed=200)
    torch.manual_seed(0)
    y, idx, hook = _run(model, "infer_panel_batch_infer", (ids, lens, prompt, bert),
                        dict(top_k=15, top_p=1.0, temperature=1.0, early_stop_num=15, repetition_penalty=1.35, max_len=max(L)))
    assert all(int(t.shape[0]) == 150 + 15 for t in y), [t.shape for t in y]  # nobody hit EOS inside the window
    emitted = np.stack([np.array([int(smp[b, 0]) for smp in hook.samples]) for b in range(32)])  # [32, 16] incl. the dropped one
    np.savez_compressed(os.path.join(GOLD, "cfg2_b32.npz"), weight_seed=0, input_seed=200, lens_seed=100, phoneme_lens=L,
                        prompt_len=150, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35, early_stop_num=15,
                        logits=_pad_logits(hook.logits), emitted=emitted,
                        y=np.stack([t.numpy() for t in y]), idx=np.array(idx))
    print("cfg2_b32: idx", idx[:4], "...", "logits", len(hook.logits))


def case_fp16w_b1():
    sd = synthetic.make_state_dict(seed=0, rounding="fp16")
    model = ref_harness.build_reference_model(sd, synthetic.S1V2_CONFIG)
    ids, lens, prompt, bert = synthetic.make_inputs(1, [80], 150, seed=1)
    y, idx, hook = _run(model, "infer_panel", (ids[0][None], lens, prompt, bert[0][None]),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=24, repetition_penalty=1.35))
    np.savez_compressed(os.path.join(GOLD, "fp16w_b1.npz"), weight_seed=0, rounding="fp16", input_seed=1, phoneme_lens=[80],
                        prompt_len=150, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35, early_stop_num=24,
                        logits=_pad_logits(hook.logits), y=y.numpy(), idx=idx)
    print("fp16w_b1: idx", idx)


def case_to_batch():
    """TTS.to_batch / TTS.recovery_order (TTS_infer_pack/TTS.py:842-973) run from the reference's OWN source: TTS.py itself cannot
    be imported here (librosa, peft, ffmpeg, ...), so the two method definitions are cut out of the file with `ast` and executed
    as plain functions (they do not touch `self`).  Recorded: the batch index lists for seeded length sets."""
    import ast
    import json
    src = open("/root/reference/GPT_SoVITS/TTS_infer_pack/TTS.py").read()
    tree = ast.parse(src)
    fns = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("to_batch", "recovery_order"):
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "np": np, "List": list}
            exec(compile(mod, "TTS.py", "exec"), ns)
            fns[node.name] = ns[node.name]
    rs = np.random.RandomState(21)
    cases = []
    for n, bs, thr, split in ((1, 5, 0.75, True), (7, 3, 0.75, True), (23, 5, 0.75, True), (120, 20, 0.75, True), (120, 32, 0.9, True),
                              (40, 8, 0.5, True), (40, 8, 1.0, True), (17, 4, 0.75, False), (64, 200, 0.75, True), (30, 1, 0.75, True)):
        lens = [int(v) for v in rs.randint(3, 140, size=n)]
        if n == 40:
            lens[:10] = [50] * 10  # ties: the sort must be stable
        data = [{"norm_text": "x" * L, "phones": [1] * (L + 2), "bert_features": torch.zeros(1024, L + 2)} for L in lens]
        batches, index_list = fns["to_batch"](None, data, prompt_data=None, batch_size=bs, threshold=thr, split_bucket=split)
        assert [len(b["all_phones"]) for b in batches] == [len(i) for i in index_list]
        rec = fns["recovery_order"](None, [[lens[i] for i in idxs] for idxs in index_list], index_list)
        assert rec == lens
        cases.append({"lengths": lens, "batch_size": bs, "threshold": thr, "split_bucket": split,
                      "batch_index_list": [[int(i) for i in idxs] for idxs in index_list]})
    json.dump(cases, open(os.path.join(GOLD, "to_batch.json"), "w"))
    print("to_batch:", len(cases), "cases;", [len(c["batch_index_list"]) for c in cases], "batches")


S1_V1_MODEL = {"vocab_size": 1025, "phoneme_vocab_size": 512, "embedding_dim": 512, "hidden_dim": 512, "head": 16,
               "linear_units": 2048, "n_layer": 12, "dropout": 0, "EOS": 1024}  # GPT_SoVITS/configs/s1.yaml (the v1 architecture)


def case_ckpt_s1v1():
    """The v1 s1 architecture (configs/s1.yaml: 12 layers, 512 phonemes) loaded the way TTS.init_t2s_weights loads a checkpoint
    (fp16 tensors under "weight" with the Lightning "model." prefix, TTS.py:585-599): the reference model receives the fp16 values
    (as fp32) and runs infer_panel_batch_infer; the GPU test writes the same checkpoint file and goes file -> engine."""
    cfg = {"model": dict(S1_V1_MODEL)}
    sd = synthetic.make_state_dict(seed=21, config=cfg, rounding="fp16")
    model = ref_harness.build_reference_model(sd, cfg)
    L = [31, 56, 44]
    ids, lens, prompt, bert = synthetic.make_inputs(3, L, 72, seed=22, phoneme_vocab=512)
    y, idx, hook = _run(model, "infer_panel_batch_infer", (ids, lens, prompt, bert),
                        dict(top_k=1, top_p=1.0, temperature=1.0, early_stop_num=20, repetition_penalty=1.35, max_len=max(L)))
    np.savez_compressed(os.path.join(GOLD, "ckpt_s1v1.npz"), weight_seed=21, rounding="fp16", input_seed=22, phoneme_lens=L,
                        prompt_len=72, phoneme_vocab=512, n_layer=12, top_k=1, repetition_penalty=1.35, early_stop_num=20,
                        logits=_pad_logits(hook.logits), y=np.stack([t.numpy() for t in y]), idx=np.array(idx))
    print("ckpt_s1v1: idx", idx)


ROUND2 = ("long_b2", "long_b1", "naive_batched", "cfg2_b32", "fp16w_b1", "to_batch", "ckpt_s1v1")


def main():
    assert ref_harness.reference_available(), "run in the build container (needs /root/reference)"
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:
        assert all(n in ROUND2 for n in only), only
        sd = synthetic.make_state_dict(seed=0)
        model = ref_harness.build_reference_model(sd, synthetic.S1V2_CONFIG)
        for n in only:
            fn = globals()["case_" + n]
            fn(model) if n in ("long_b2", "long_b1", "cfg2_b32") else fn()
        return
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    sd = synthetic.make_state_dict(seed=0)
    model = ref_harness.build_reference_model(sd, synthetic.S1V2_CONFIG)
    case_naive_b1(model)
    case_batch_b4(model)
    case_reffree_b1(model)
    case_retire_b6()
    case_sampler_kat()
    case_latent()
    case_long_b2(model)
    case_long_b1(model)
    case_naive_batched()
    case_cfg2_b32(model)
    case_fp16w_b1()
    case_to_batch()
    case_ckpt_s1v1()


if __name__ == "__main__":
    main()
