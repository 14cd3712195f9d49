"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).

The CUDA path is called through the C-ABI (gpt_sovits_b200.T2SEngine -> libt2s_b200.so) and compared
with (a) the golden vectors recorded from the live reference (tests/golden/*.npz) and (b) the numpy
oracle (oracle/) on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  * teacher-forced per-step logits: |delta| <= LOGIT_TOL.  The engine keeps weights, GEMM operands and the
    KV cache in bf16 with fp32 accumulation / residual / LayerNorm / softmax; the reference is fp32 on the
    same bf16-representable weights.  Emulating exactly that recipe inside the reference gave a max
    error of 0.028 at logit sigma~1 -> LOGIT_TOL = 0.06 (2x).
  * greedy tokens: bit-exact wherever the reference's top-1 margin exceeds 2*LOGIT_TOL.
  * seeded sampling: bit-exact against the fp32 replay (oracle/sampler_oracle.py) of the engine's own
    captured logits, wherever the exponential-race margin exceeds RACE_TOL (expf/logf on the device and
    numpy's libm may differ in the last ulp).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from gpt_sovits_b200 import synthetic

pytestmark = pytest.mark.gpu

LOGIT_TOL = 0.06
RACE_TOL = 1e-4


@pytest.fixture(scope="module")
def engine(weights_seed0, pe_table):
    from gpt_sovits_b200 import T2SEngine
    eng = T2SEngine(synthetic.S1V2_CONFIG, device="cuda:0")
    eng.load_state_dict(weights_seed0, pe=pe_table)
    yield eng
    eng.close()


@pytest.fixture(autouse=True)
def _auto_decode_mode(request):
    """Every test starts from the product default: decode mode 5 = auto (the cluster-stream kernel, mode 4, whenever the
    batch fits its 16-CTA clusters; the grid-wide phase kernels otherwise)."""
    if "engine" in request.fixturenames:
        from gpt_sovits_b200 import _lib
        request.getfixturevalue("engine").set_option(_lib.OPT_DECODE_MODE, 5)
    yield


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _inputs(g, device="cuda:0"):
    L = [int(v) for v in g["phoneme_lens"]]
    ids, lens, prompt, bert = synthetic.make_inputs(len(L), L, int(g["prompt_len"]), seed=int(g["input_seed"]))
    ids = [t.to(device) for t in ids]
    bert = [t.to(device) for t in bert]
    prompt = None if prompt is None else prompt.to(device)
    return ids, bert, prompt


def _compare_logits(res, g, eos_window, n_steps, slots_per_step=None):
    """Returns (max abs error, greedy agreements, greedy comparisons made)."""
    ref = g["logits"]
    got = res.logits.cpu().numpy()  # [n, B, 1025] by slot
    worst, agree, total = 0.0, 0, 0
    for s in range(n_steps):
        width = 1024 if s < eos_window else 1025
        slots = slots_per_step[s] if slots_per_step is not None else list(range(got.shape[1]))
        for r, b in enumerate(slots):
            rr = ref[s, r, :width]
            assert not np.isnan(rr).any()
            gg = got[s, b, :width]
            assert not np.isnan(gg).any(), f"step {s} slot {b}: logits were not produced"
            worst = max(worst, float(np.abs(gg - rr).max()))
            top2 = np.partition(rr, -2)[-2:]
            if top2[1] - top2[0] > 2 * LOGIT_TOL:  # raw-logit margin; greedy cases use rp on both sides
                total += 1
                agree += int(np.argmax(gg) == np.argmax(rr))
    return worst, agree, total


@pytest.mark.parametrize("mode", [0, 1, 4])
def test_naive_b1_teacher_forced(engine, golden_dir, mode):
    """Config-1 shape (B=1, 80 phonemes + 150 prompt), reference infer_panel_naive goldens."""
    from gpt_sovits_b200 import _lib
    engine.set_option(_lib.OPT_DECODE_MODE, mode)
    g = _golden(golden_dir, "naive_b1")
    ids, bert, prompt = _inputs(g)
    P, idx = int(g["prompt_len"]), int(g["idx"])
    forced = torch.from_numpy(g["y"][:, P:]).to(torch.int32)
    n = g["logits"].shape[0]
    res = engine.infer(ids, bert, prompt, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                       early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=11, forced=forced,
                       capture_logits=n)
    worst, agree, total = _compare_logits(res, g, 11, n)
    print(f"naive_b1 mode={mode}: max |dlogit| = {worst:.4f}, greedy agree {agree}/{total}")
    assert worst <= LOGIT_TOL
    assert agree == total
    assert res.idx == [idx]
    np.testing.assert_array_equal(res.sequences()[0].cpu().numpy(), g["y"][0])


def test_naive_b1_free_running(engine, golden_dir):
    g = _golden(golden_dir, "naive_b1")
    ids, bert, prompt = _inputs(g)
    res = engine.infer(ids, bert, prompt, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                       early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=11)
    assert res.idx == [int(g["idx"])]
    got = res.sequences()[0].cpu().numpy()
    ref = g["y"][0]
    if not np.array_equal(got, ref):
        # only acceptable where the reference's own top-1 margin is inside the tolerance
        first = int(np.nonzero(got != ref)[0][0]) - int(g["prompt_len"])
        row = g["logits"][first, 0, : (1024 if first < 11 else 1025)]
        top2 = np.partition(row, -2)[-2:]
        assert top2[1] - top2[0] <= 2 * LOGIT_TOL, f"greedy diverged at step {first} with margin {top2[1]-top2[0]:.3f}"


@pytest.mark.parametrize("mode", [0, 1, 4])
def test_batch_b4_teacher_forced(engine, golden_dir, mode):
    """Ragged batch through infer_panel_batch_infer semantics (EOS column dropped at idx 0 only)."""
    from gpt_sovits_b200 import _lib
    engine.set_option(_lib.OPT_DECODE_MODE, mode)
    g = _golden(golden_dir, "batch_b4")
    ids, bert, prompt = _inputs(g)
    P = int(g["prompt_len"])
    forced = torch.from_numpy(g["y"][:, P:]).to(torch.int32)
    n = g["logits"].shape[0]
    res = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=1,
                       forced=forced, capture_logits=n)
    worst, agree, total = _compare_logits(res, g, 1, n)
    print(f"batch_b4 mode={mode}: max |dlogit| = {worst:.4f}, greedy agree {agree}/{total}")
    assert worst <= LOGIT_TOL and agree == total
    assert res.idx == [int(v) for v in g["idx"]]
    for b in range(4):
        np.testing.assert_array_equal(res.sequences()[b].cpu().numpy(), g["y"][b])


@pytest.mark.parametrize("mode", [0, 1, 4])
def test_retirement_b6(golden_dir, pe_table, mode):
    """EOS-prone head: sequences retire at different steps (on-device compaction); outputs must come
    back in the original order with the reference's idx (t2s_model.py:724-745,779)."""
    from gpt_sovits_b200 import T2SEngine, _lib
    g = _golden(golden_dir, "retire_b6")
    sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), eos_scale=float(g["eos_scale"]))
    eng = T2SEngine(synthetic.S1V2_CONFIG, device="cuda:0")
    try:
        eng.load_state_dict(sd, pe=pe_table)
        eng.set_option(_lib.OPT_DECODE_MODE, mode)
        ids, bert, prompt = _inputs(g)
        P = int(g["prompt_len"])
        ref_idx = [int(v) for v in g["idx"]]
        # reconstruct the active list per step from the reference's idx
        n = g["logits"].shape[0]
        slots_per_step = [[b for b in range(6) if ref_idx[b] >= s] for s in range(n)]
        # The golden has several top-1 margins below 0.01 (near-ties flip under bf16 noise and change the
        # trajectory), so the run is teacher-forced with the reference's tokens, and with EOS at each
        # sequence's stopping step (greedy: the reference sampled its argmax = EOS there): retirement then
        # happens at exactly the reference's steps and every step's logits are comparable.
        forced = torch.full((6, n), 0, dtype=torch.int32)
        for b in range(6):
            y = g["y"][b]
            y = y[y >= 0][P:]
            forced[b, : len(y)] = torch.from_numpy(y).to(torch.int32)
            forced[b, ref_idx[b]] = 1024
        res = eng.infer(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=1,
                        forced=forced, capture_logits=n)
        print(f"retire_b6 mode={mode}: idx {res.idx} (ref {ref_idx})")
        assert res.idx == ref_idx
        worst, agree, total = _compare_logits(res, g, 1, n, slots_per_step)
        print(f"retire_b6 mode={mode}: max |dlogit| = {worst:.4f}, greedy agree {agree}/{total}")
        assert worst <= LOGIT_TOL and agree == total
        for b in range(6):
            y = g["y"][b]
            np.testing.assert_array_equal(res.sequences()[b].cpu().numpy(), y[y >= 0])
            # the engine's own (un-forced) choice at the stopping step is EOS too where the margin allows
            s_stop = ref_idx[b]
            row = g["logits"][s_stop, slots_per_step[s_stop].index(b), :1025].copy()
            from oracle import sampler_oracle as so
            so.apply_repetition_penalty(row, y[: P + s_stop], 1.35)
            top2 = np.sort(row)[-2:]
            if top2[1] - top2[0] > 2 * LOGIT_TOL:
                assert int(res.sampled[b, s_stop]) == 1024
    finally:
        eng.close()


def test_reference_free(engine, golden_dir):
    """prompts=None: no prompt rows, empty penalty history, idx reported as 0 (t2s_model.py:849-856,:916)."""
    import gpt_sovits_b200 as gsb
    g = _golden(golden_dir, "reffree_b1")
    ids, bert, _ = _inputs(g)
    forced = torch.from_numpy(g["y"]).to(torch.int32)  # teacher-forced: the golden has near-tie steps
    res = engine.infer(ids, bert, None, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=11,
                       forced=forced, capture_logits=g["logits"].shape[0])
    worst, agree, total = _compare_logits(res, g, 11, g["logits"].shape[0])
    print(f"reffree: max |dlogit| = {worst:.4f}, greedy agree {agree}/{total}")
    assert worst <= LOGIT_TOL and agree == total
    np.testing.assert_array_equal(res.sequences()[0].cpu().numpy(), g["y"][0])


def test_sampler_kernel_known_answers(engine, golden_dir):
    """The fused sampling kernel alone on the reference's logits_to_probs rows: the sampled token must be
    the fp32 replay's argmax(probs/q) and the greedy id the argmax of the reference-penalised logits."""
    from gpt_sovits_b200 import _lib
    from oracle import sampler_oracle as so
    g = _golden(golden_dir, "sampler_kat")
    lib = engine.lib
    seed, step = 0x1234_5678_9ABC, 7
    checked = 0
    for i in range(g["logits"].shape[0]):
        w = int(g["width"][i])
        if w not in (1024, 1025):
            continue  # the 4-wide hand rows are covered on the CPU side
        temperature, top_k, top_p, rp = [float(v) for v in g["params"][i]]
        prev = g["prev"][i]
        prev = prev[prev >= 0].astype(np.int32)
        row = np.zeros((1, 1025), np.float32)
        row[0, :w] = g["logits"][i, :w]
        tok, greedy = (C.c_int32 * 1)(), (C.c_int32 * 1)()
        pv = np.ascontiguousarray(prev if prev.size else np.array([-1], np.int32))
        _lib.check(lib.t2s_sampler_test(engine._h, row.ctypes.data, 1, w, pv.ctypes.data, int(pv.size), int(top_k),
                                        top_p, temperature, rp, C.c_uint64(seed), step, tok, greedy, None))
        q = so.exp_noise(seed, 0, step, w)
        lg = g["logits"][i, :w].copy()
        t_ref, g_ref, probs, margin = so.sample_row(lg, prev, q, temperature, int(top_k), top_p, rp)
        assert greedy[0] == int(np.argmax(g["penalised"][i, :w])) == g_ref
        ref_support = g["probs"][i, :w] > 0
        assert ref_support[tok[0]], f"row {i}: sampled token {tok[0]} is outside the reference's support"
        if margin > RACE_TOL:
            assert tok[0] == t_ref, f"row {i}: {tok[0]} != replay {t_ref} (margin {margin:.2e})"
        checked += 1
    assert checked >= 6


def test_seeded_sampling_replay(engine, golden_dir):
    """Config-2 sampling parameters (top_k=15, top_p=1, T=1, rp=1.35) with a fixed seed: every sampled
    token must equal the fp32 host replay of the engine's own logits with the same Philox stream."""
    from oracle import sampler_oracle as so
    g = _golden(golden_dir, "batch_b4")
    ids, bert, prompt = _inputs(g)
    P, n, seed = int(g["prompt_len"]), 24, 20240607
    res = engine.infer(ids, bert, prompt, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                       early_stop_num=n - 1, eos_suppress_steps=1, seed=seed, capture_logits=n)
    logits = res.logits.cpu().numpy()
    seqs = [s.cpu().numpy() for s in res.sequences()]
    mismatches, gated = 0, 0
    for b in range(4):
        hist = list(map(int, prompt[b].cpu().numpy()))
        for s in range(min(n, res.idx[b] + 1)):
            w = 1024 if s < 1 else 1025
            row = logits[s, b, :w].copy()
            assert not np.isnan(row).any()
            q = so.exp_noise(seed, b, s, w)
            tok, greedy, _, margin = so.sample_row(row, hist, q, 1.0, 15, 1.0, 1.35)
            got = int(res.sampled[b, s])
            if margin > RACE_TOL:
                mismatches += int(got != tok)
            else:
                gated += 1
            hist.append(got)
    print(f"sampling replay: {mismatches} mismatches, {gated} near-ties skipped")
    assert mismatches == 0
    assert gated <= 2
    # sampled tokens are what the engine emitted
    for b in range(4):
        np.testing.assert_array_equal(seqs[b][P:], res.sampled[b, : res.idx[b]].numpy())


def test_top_p_and_temperature_path(engine, golden_dir):
    """top_p < 1 and temperature != 1 (bitonic-sort path), replayed on the host."""
    from oracle import sampler_oracle as so
    g = _golden(golden_dir, "naive_b1")
    ids, bert, prompt = _inputs(g)
    n, seed = 12, 99
    res = engine.infer(ids, bert, prompt, top_k=20, top_p=0.85, temperature=0.8, repetition_penalty=1.2,
                       early_stop_num=n - 1, eos_suppress_steps=11, seed=seed, capture_logits=n)
    logits = res.logits.cpu().numpy()
    hist = list(map(int, prompt[0].cpu().numpy()))
    bad = 0
    for s in range(min(n, res.idx[0] + 1)):
        row = logits[s, 0, :1024].copy()
        q = so.exp_noise(seed, 0, s, 1024)
        tok, _, probs, margin = so.sample_row(row, hist, q, 0.8, 20, 0.85, 1.2)
        got = int(res.sampled[0, s])
        assert probs[got] > 0 or margin <= RACE_TOL, f"step {s}: token {got} outside the top-p/top-k support"
        if margin > RACE_TOL:
            bad += int(got != tok)
        hist.append(got)
    assert bad == 0


def test_against_numpy_oracle_and_batch_invariance(engine, weights_seed0, pe_table):
    """Fresh seeded inputs (no golden): oracle vs CUDA teacher-forced, plus the ragged-batch invariant the
    reference has: an utterance's logits do not depend on its batch."""
    from oracle.t2s_oracle import T2SOracle
    L = [37, 64, 65, 20, 128]
    idsc, lens, promptc, bertc = synthetic.make_inputs(5, L, 77, seed=11)
    o = T2SOracle(weights_seed0, pe_table)
    n = 10
    out = o.generate([t.numpy() for t in idsc], [t.numpy() for t in bertc], promptc.numpy(), top_k=1,
                     early_stop_num=n - 1, eos_window=1, record_logits=True)
    forced = torch.tensor([out["generated"][b][:n] for b in range(5)], dtype=torch.int32)
    ids = [t.cuda() for t in idsc]
    bert = [t.cuda() for t in bertc]
    prompt = promptc.cuda()
    res = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced,
                       capture_logits=n)
    got = res.logits.cpu().numpy()
    worst = 0.0
    for s in range(n):
        w = 1024 if s < 1 else 1025
        worst = max(worst, float(np.abs(got[s, :, :w] - out["logits"][s][:, :w]).max()))
    print(f"oracle vs CUDA: max |dlogit| = {worst:.4f}")
    assert worst <= LOGIT_TOL
    # utterance 2 alone, same forced tokens
    res1 = engine.infer([ids[2]], [bert[2]], prompt[2:3], top_k=1, early_stop_num=n - 1, eos_suppress_steps=1,
                        forced=forced[2:3], capture_logits=n)
    d = float(np.abs(res1.logits.cpu().numpy()[:, 0, :1024] - got[:, 2, :1024]).max())
    print(f"batch invariance: max |dlogit| = {d:.5f}")
    # Not bit-identical: the split-KV partition (hence the fp32 summation order) depends on the batch, a
    # 1-ulp difference can flip a bf16 rounding, and this deliberately sensitive 24-layer post-LN init
    # amplifies one flip ~500x (the reference's own fp32 noise floor of 4.5e-5 is amplified the same way,
    # SURVEY.md section 8a item 3).  The bound is therefore the logits tolerance itself.
    assert d <= LOGIT_TOL


def test_bitwise_reproducible_and_modes_identical(engine, golden_dir):
    """Default (deterministic) arithmetic has no floating-point atomics: the same request twice gives
    bit-identical logits, and the persistent cooperative kernel (grid barriers) gives bit-identical logits
    to the one-kernel-per-phase graph path -- any missed barrier / stale-L1 read would show up here."""
    from gpt_sovits_b200 import _lib
    g = _golden(golden_dir, "retire_b6")  # shapes only; seed-0 weights: sequences retire at other steps
    ids, bert, prompt = _inputs(g)
    n = 40
    outs = []
    def run(mode):
        engine.set_option(_lib.OPT_DECODE_MODE, mode)
        r = engine.infer(ids, bert, prompt, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                         early_stop_num=n - 1, eos_suppress_steps=1, seed=4242, capture_logits=n)
        assert int(r.stats["decode_mode"]) == (4 if mode == 5 else mode)
        return r.logits.cpu().numpy(), r.tokens.cpu().numpy(), r.idx
    for mode in (0, 1, 1, 0):
        outs.append(run(mode))
    for lg, tk, ix in outs[1:]:
        assert ix == outs[0][2]
        np.testing.assert_array_equal(tk, outs[0][1])
        np.testing.assert_array_equal(np.nan_to_num(lg, nan=-7.0), np.nan_to_num(outs[0][0], nan=-7.0))
    # the cluster-stream kernel (mode 4 = what auto picks for this batch) sums in a different order than the phase kernels,
    # so it is compared with itself: cluster barriers / DSMEM hand-offs / the TMA ring must be race free
    cs = [run(4), run(5), run(4)]
    for lg, tk, ix in cs[1:]:
        assert ix == cs[0][2]
        np.testing.assert_array_equal(tk, cs[0][1])
        np.testing.assert_array_equal(np.nan_to_num(lg, nan=-7.0), np.nan_to_num(cs[0][0], nan=-7.0))
    d = float(np.nanmax(np.abs(cs[0][0][:, :, :1024] - outs[0][0][:, :, :1024])))
    print(f"cluster-stream vs phase kernels (free-running sampled run, first divergence amplifies): max |dlogit| over steps = {d:.4f}")


@pytest.mark.parametrize("gemm", [1, 2])
@pytest.mark.parametrize("case,window", [("naive_b1", 11), ("batch_b4", 1)])
def test_prefill_tcgen05_gemm(engine, golden_dir, case, window, gemm):
    """Prefill with the tcgen05/TMEM + TMA GEMMs (T2S_OPT_PREFILL_GEMM = 1: the persistent 128 x 256-tile kernel, the default;
    2: round 1's one-tile-per-CTA kernel): teacher-forced logits within the tolerance of the reference goldens, and close to
    the warp-MMA prefill path (different summation order)."""
    from gpt_sovits_b200 import _lib
    g = _golden(golden_dir, case)
    ids, bert, prompt = _inputs(g)
    P = int(g["prompt_len"])
    forced = torch.from_numpy(g["y"][:, P:]).to(torch.int32)
    n = g["logits"].shape[0]
    kw = dict(top_k=1, early_stop_num=int(g["early_stop_num"]), eos_suppress_steps=window, forced=forced, capture_logits=n)
    try:
        engine.set_option(_lib.OPT_PREFILL_GEMM, gemm)
        res = engine.infer(ids, bert, prompt, **kw)
        worst, agree, total = _compare_logits(res, g, window, n)
        engine.set_option(_lib.OPT_PREFILL_GEMM, 0)
        ref = engine.infer(ids, bert, prompt, **kw)
    finally:
        engine.set_option(_lib.OPT_PREFILL_GEMM, 1)
    d = float(np.nanmax(np.abs(res.logits.cpu().numpy()[:, :, :1024] - ref.logits.cpu().numpy()[:, :, :1024])))
    print(f"tcgen05 prefill (gemm {gemm}) {case}: max |dlogit| vs reference = {worst:.4f}, vs warp-MMA prefill = {d:.4f}, greedy {agree}/{total}")
    assert worst <= LOGIT_TOL and agree == total
    assert d <= LOGIT_TOL


def test_large_batch_tcgen05_decode(engine):
    """Large batches (default >= 160; forced to 64 here) switch the decode projections to the tcgen05 GEMMs (CUDA graph, mode 3).  Same teacher-forced
    run with the persistent warp-MMA kernel (mode 1 forced): logits agree within the tolerance, and the
    retirement bookkeeping (EOS forced at different steps) gives identical idx and tokens."""
    from gpt_sovits_b200 import _lib
    B, P, n = 72, 30, 10
    L = synthetic.config_lens(B, 16, 48, seed=9)
    ids, lens, prompt, bert = synthetic.make_inputs(B, L, P, seed=31)
    ids = [t.cuda() for t in ids]
    bert = [t.cuda() for t in bert]
    prompt = prompt.cuda()
    g = torch.Generator().manual_seed(3)
    forced = torch.randint(0, 1024, (B, n), dtype=torch.int32, generator=g)
    stop = torch.randint(3, n, (B,), generator=g)
    for b in range(0, B, 3):
        forced[b, int(stop[b])] = 1024  # every third utterance retires early
    kw = dict(top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced, capture_logits=n)
    try:
        engine.set_option(_lib.OPT_TC_DECODE_MIN_BATCH, 64)  # default crossover is 160
        res_tc = engine.infer(ids, bert, prompt, **kw)
        assert int(res_tc.stats["decode_mode"]) == 3
        engine.set_option(_lib.OPT_TC_DECODE_MIN_BATCH, 0)
        res_pk = engine.infer(ids, bert, prompt, **kw)
        assert int(res_pk.stats["decode_mode"]) == 1
    finally:
        engine.set_option(_lib.OPT_TC_DECODE_MIN_BATCH, 160)
    assert res_tc.idx == res_pk.idx
    assert len(set(res_tc.idx)) > 3
    a, b_ = res_tc.logits.cpu().numpy(), res_pk.logits.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b_))
    d = float(np.nanmax(np.abs(a[:, :, :1024] - b_[:, :, :1024])))
    print(f"tcgen05 decode (B={B}) vs persistent warp-MMA: max |dlogit| = {d:.4f}")
    assert d <= LOGIT_TOL
    assert torch.equal(res_tc.tokens, res_pk.tokens)
    assert int(res_tc.stats["decode_steps"]) == int(res_pk.stats["decode_steps"]) == max(res_tc.idx)


@pytest.mark.parametrize("B,P,lo,hi", [(20, 30, 16, 48), (56, 140, 100, 260), (3, 1900, 380, 400)])
def test_cluster_stream_multi_row_vs_phase_kernels(engine, B, P, lo, hi):
    """The cluster-stream kernel with several sequences per cluster (3 and 8 rows per 16-CTA cluster; the second case also
    crosses 128-position K/V page boundaries: up to 3 pages per sequence, partial last pages; the third has 18 pages per
    sequence: several pages per warp and ring slot) against the grid-wide phase
    kernels on the same teacher-forced run with EOS forced at different steps: logits within the tolerance, identical
    retirement (idx) and tokens."""
    from gpt_sovits_b200 import _lib
    n = 12
    L = synthetic.config_lens(B, lo, hi, seed=19)
    ids, lens, prompt, bert = synthetic.make_inputs(B, L, P, seed=37)
    ids = [t.cuda() for t in ids]
    bert = [t.cuda() for t in bert]
    prompt = prompt.cuda()
    g = torch.Generator().manual_seed(5)
    forced = torch.randint(0, 1024, (B, n), dtype=torch.int32, generator=g)
    stop = torch.randint(3, n, (B,), generator=g)
    for b in range(0, B, 3):
        forced[b, int(stop[b])] = 1024  # every third utterance retires early: rows are re-dealt to the clusters
    kw = dict(top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced, capture_logits=n)
    engine.set_option(_lib.OPT_DECODE_MODE, 5)
    res_cs = engine.infer(ids, bert, prompt, **kw)
    assert int(res_cs.stats["decode_mode"]) == 4
    engine.set_option(_lib.OPT_DECODE_MODE, 1)
    res_pk = engine.infer(ids, bert, prompt, **kw)
    assert int(res_pk.stats["decode_mode"]) == 1
    assert res_cs.idx == res_pk.idx
    assert len(set(res_cs.idx)) > min(3, B - 2)
    a, b_ = res_cs.logits.cpu().numpy(), res_pk.logits.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b_))
    d = float(np.nanmax(np.abs(a[:, :, :1024] - b_[:, :, :1024])))
    print(f"cluster-stream (B={B}) vs phase kernels: max |dlogit| = {d:.4f}")
    assert d <= LOGIT_TOL
    assert torch.equal(res_cs.tokens, res_pk.tokens)


def test_large_batch_runs_as_cluster_stream_chunks(engine):
    """Auto mode: a batch that does not fit the cluster-stream kernel (56 sequences) and is below the tcgen05 crossover is
    run by t2s_generate as equal chunks that do fit.  Greedy decoding is deterministic, so the chunked call must give exactly
    what the two halves give when they are submitted on their own."""
    B, P, n = 72, 30, 8
    L = synthetic.config_lens(B, 16, 48, seed=23)
    ids, lens, prompt, bert = synthetic.make_inputs(B, L, P, seed=41)
    ids = [t.cuda() for t in ids]
    bert = [t.cuda() for t in bert]
    prompt = prompt.cuda()
    kw = dict(top_k=1, early_stop_num=n, eos_suppress_steps=1)
    res = engine.infer(ids, bert, prompt, **kw)
    assert int(res.stats["decode_mode"]) == 4
    assert int(res.stats["decode_steps"]) == 2 * n  # two chunks, n decode steps each (prefill samples step 0)
    h = B // 2
    ra = engine.infer(ids[:h], bert[:h], prompt[:h], **kw)
    rb = engine.infer(ids[h:], bert[h:], prompt[h:], **kw)
    assert res.idx == ra.idx + rb.idx
    assert torch.equal(res.tokens[:h], ra.tokens) and torch.equal(res.tokens[h:], rb.tokens)
    # sampling: every utterance keeps its own Philox stream (keyed by its index in the caller's batch), chunked or not
    kw = dict(top_k=15, temperature=1.0, early_stop_num=n, eos_suppress_steps=1, seed=77)
    r1 = engine.infer(ids, bert, prompt, **kw)
    r2 = engine.infer(ids, bert, prompt, **kw)
    assert r1.idx == r2.idx and torch.equal(r1.tokens, r2.tokens)
    assert len({tuple(r1.tokens[b, P:P + n].tolist()) for b in range(B)}) > B // 2


def test_engine_from_checkpoint(tmp_path, pe_table):
    """Checkpoint file (the reference's fp16 {"weight","config","info"} format) -> engine gives the same greedy tokens as the
    same weights handed over as a state_dict."""
    import gpt_sovits_b200 as gsb
    cfg = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=4)}
    sd = {k: v.half().float() for k, v in synthetic.make_state_dict(seed=11, config=cfg).items()}  # fp16-representable
    path = str(tmp_path / "s1.ckpt")
    torch.save({"weight": {"model." + k: v.half() for k, v in sd.items()}, "config": dict(cfg, data={"max_sec": 54}), "info": "t"}, path)
    ids, lens, prompt, bert = synthetic.make_inputs(3, [20, 31, 26], 24, seed=2)
    ids, bert, prompt = [t.cuda() for t in ids], [t.cuda() for t in bert], prompt.cuda()
    kw = dict(top_k=1, early_stop_num=10, eos_suppress_steps=1)
    eng, config = gsb.engine_from_checkpoint(path, pe=pe_table)
    try:
        assert config["data"]["max_sec"] == 54
        a = eng.infer(ids, bert, prompt, **kw)
    finally:
        eng.close()
    ref = gsb.T2SEngine(cfg)
    try:
        ref.load_state_dict(sd, pe=pe_table)
        b = ref.infer(ids, bert, prompt, **kw)
    finally:
        ref.close()
    assert a.idx == b.idx and torch.equal(a.tokens, b.tokens)


def test_streaming_decode_in_slices(golden_dir, pe_table):
    """The streaming form (t2s_decode with a step budget + t2s_result in between): sequences that stop early are reported
    while the others keep decoding, every relaunch of the cluster-stream kernel picks the session up where it stopped, and
    the final result equals the one-shot call.  (EOS-prone weights: sequences retire at different steps.)"""
    from gpt_sovits_b200 import T2SEngine
    g = _golden(golden_dir, "retire_b6")
    sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), eos_scale=float(g["eos_scale"]))
    eng = T2SEngine(synthetic.S1V2_CONFIG, device="cuda:0")
    try:
        eng.load_state_dict(sd, pe=pe_table)
        ids, bert, prompt = _inputs(g)
        P = int(g["prompt_len"])
        kw = dict(top_k=1, early_stop_num=60, eos_suppress_steps=1)
        whole = eng.infer(ids, bert, prompt, **kw)
        assert int(whole.stats["decode_mode"]) == 4
        part = eng.infer(ids, bert, prompt, max_new_steps=7, **kw)
        seen_done = [i >= 0 for i in part.idx]
        slices = 1
        while not all(i >= 0 for i in part.idx):
            assert eng.decode_more(7) > 0
            part = eng.result(len(ids), P)
            for b, i in enumerate(part.idx):
                assert not (seen_done[b] and i < 0)  # finished stays finished
                seen_done[b] = seen_done[b] or i >= 0
            slices += 1
            assert slices < 40
        assert slices > 2
        assert part.idx == whole.idx
        for b in range(len(ids)):
            n = P + max(whole.idx[b], 0)
            assert torch.equal(part.tokens[b, :n], whole.tokens[b, :n])
    finally:
        eng.close()


def test_drop_in_patch_with_fake_tts_caller(weights_seed0, pe_table):
    """The class-level patch, driven the way TTS.run drives the reference (TTS.py:1042-1047, 1210-1227,
    1259): instance-level rebinding to the batched variant, prompt as an .expand view, fp16 BERT
    features, results sliced with [-idx:]."""
    import gpt_sovits_b200 as gsb

    class FakeDecoder(torch.nn.Module):
        """Attribute-compatible stand-in for Text2SemanticDecoder (the real class needs /root/reference)."""

        def __init__(self, sd, pe):
            super().__init__()
            self.model_dim = self.embedding_dim = 512
            self.num_head, self.num_layers, self.vocab_size, self.phoneme_vocab_size, self.EOS = 16, 24, 1025, 732, 1024
            self._params = torch.nn.ParameterDict()
            self._sd_keys = {}
            for i, (k, v) in enumerate(sd.items()):
                name = f"p{i}"
                self._params[name] = torch.nn.Parameter(v.clone(), requires_grad=False)
                self._sd_keys[name] = k
            self.ar_audio_position = type("PE", (), {"pe": pe[None]})()

        def state_dict(self, *a, **k):
            return {self._sd_keys[n]: p for n, p in self._params.items()}

    gsb.patch_reference(FakeDecoder)
    try:
        model = FakeDecoder(weights_seed0, pe_table).cuda()
        L = [30, 41, 25]
        ids, lens, prompt, bert = synthetic.make_inputs(3, L, 50, seed=5)
        ids = [t.cuda() for t in ids]
        bert = [t.cuda().half() for t in bert]
        prompt = prompt[0:1].cuda().expand(3, -1)
        assert prompt.stride(0) == 0
        model.infer_panel = model.infer_panel_batch_infer  # what TTS.run does per request
        torch.manual_seed(123)
        y_list, idx_list = model.infer_panel(ids, lens.cuda(), prompt, bert, top_k=5, top_p=1, temperature=1.0,
                                             early_stop_num=20, max_len=41, repetition_penalty=1.35)
        assert len(y_list) == 3 and len(idx_list) == 3
        for y, idx in zip(y_list, idx_list):
            assert y.dtype == torch.int64 and y.is_cuda and y.shape[0] == 50 + idx and 0 < idx <= 20
            assert torch.equal(y[:50], prompt[0])
            tail = y[-idx:]
            assert int(tail.min()) >= 0 and int(tail.max()) < 1024
        torch.manual_seed(123)
        y2, idx2 = model.infer_panel(ids, lens.cuda(), prompt, bert, top_k=5, top_p=1, temperature=1.0,
                                     early_stop_num=20, max_len=41, repetition_penalty=1.35)
        assert idx2 == idx_list  # seed from torch's generator => reproducible request
        # single-utterance form (api.py:935): x [1,L], bert [1,1024,L]
        y, idx = model.infer_panel_naive(ids[0][None], lens[:1].cuda(), prompt[:1], bert[0][None], top_k=5, top_p=1,
                                         temperature=1.0, early_stop_num=15)
        assert y.shape == (1, 50 + idx) and idx == 15
        # weight change invalidates the packed cache
        e1 = gsb.engine_for(model)
        model.half()
        e2 = gsb.engine_for(model)
        assert e1 is not e2
        with pytest.raises(RuntimeError):
            model.infer_panel_naive(ids[0][None], lens[:1], prompt[:1], bert[0][None], top_k=0)
    finally:
        gsb.unpatch_reference(FakeDecoder)


def test_no_cpu_fallback():
    import gpt_sovits_b200 as gsb
    with pytest.raises(RuntimeError):
        gsb.T2SEngine(synthetic.S1V2_CONFIG, device="cpu")


def test_full_size_properties(engine):
    """BASELINE.json config 2 at full size (B=32, top_k=15, rp=1.35) for 200 steps: size-independent
    properties -- token range, idx == early_stop_num without EOS, prompt echoed, stats consistent with the
    KV-length arithmetic (sum of attended positions), permutation equivariance of greedy decoding."""
    B, P, n = 32, 150, 200
    L = synthetic.config_lens(B, 60, 120, seed=2)
    ids, lens, prompt, bert = synthetic.make_inputs(B, L, P, seed=21)
    ids = [t.cuda() for t in ids]
    bert = [t.cuda() for t in bert]
    prompt = prompt.cuda()
    res = engine.infer(ids, bert, prompt, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                       early_stop_num=n, eos_suppress_steps=1, seed=7)
    toks = res.tokens.cpu().numpy()
    for b in range(B):
        assert 0 <= res.idx[b] <= n
        seq = toks[b, : P + res.idx[b]]
        assert np.array_equal(seq[:P], prompt[b].cpu().numpy())
        assert seq[P:].min() >= 0 and seq[P:].max() <= 1023
        assert (toks[b, P + res.idx[b]:] == -1).all()
    st = res.stats
    # decode step s (1-based) of an active sequence attends L_b + P + s positions
    steps = int(st["decode_steps"])
    assert steps == max(res.idx)
    expect = sum(sum(L[b] + P + s for s in range(1, res.idx[b] + 1)) for b in range(B))
    assert int(st["decode_kv_positions"]) == expect
    # greedy + reversed batch order == reversed outputs (sequences never interact).  A different batch
    # composition changes the split-KV partition, so a token may flip only at a near-tie step.
    n2 = 30
    r1 = engine.infer(ids[:8], bert[:8], prompt[:8], top_k=1, early_stop_num=n2, eos_suppress_steps=1,
                      capture_logits=n2 + 1)
    r2 = engine.infer(ids[:8][::-1], bert[:8][::-1], prompt[:8], top_k=1, early_stop_num=n2, eos_suppress_steps=1)
    lg = r1.logits.cpu().numpy()
    for b, (a, c) in enumerate(zip(r1.sequences(), r2.sequences()[::-1])):
        a, c = a.cpu().numpy(), c.cpu().numpy()
        if not np.array_equal(a, c):
            s = int(np.nonzero(a != c)[0][0]) - P
            row = lg[s, b, : (1024 if s < 1 else 1025)].copy()
            seen = np.unique(a[: P + s])
            row[seen] = np.where(row[seen] < 0, row[seen] * 1.35, row[seen] / 1.35)
            top2 = np.sort(row)[-2:]
            assert top2[1] - top2[0] <= LOGIT_TOL, f"utterance {b} diverged at step {s} with margin {top2[1]-top2[0]:.3f}"


def test_edge_cases_small_and_limits(engine, weights_seed0, pe_table):
    """Minimum sizes and parameter limits the reference's UIs can reach: one phoneme,
    one prompt token; early_stop_num 0; the step cap; top_k above the vocabulary; temperature 0 (clamped to
    1e-5, effectively greedy); top_p 0 (only the arg-max survives); repetition penalty off; EOS at step 0."""
    from oracle.t2s_oracle import T2SOracle
    o = T2SOracle(weights_seed0, pe_table)
    idsc, lens, promptc, bertc = synthetic.make_inputs(2, [1, 3], 1, seed=40)
    ids = [t.cuda() for t in idsc]
    bert = [t.cuda() for t in bertc]
    prompt = promptc.cuda()
    # (a) tiny inputs against the oracle, teacher-forced
    n = 4
    out = o.generate([t.numpy() for t in idsc], [t.numpy() for t in bertc], promptc.numpy(), top_k=1,
                     early_stop_num=n - 1, eos_window=1, record_logits=True)
    forced = torch.tensor([out["generated"][b][:n] for b in range(2)], dtype=torch.int32)
    res = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=n - 1, eos_suppress_steps=1, forced=forced,
                       capture_logits=n)
    got = res.logits.cpu().numpy()
    worst = max(float(np.abs(got[s, :, :1024] - out["logits"][s][:, :1024]).max()) for s in range(n))
    assert worst <= LOGIT_TOL, worst
    assert res.idx == out["idx"]
    # (b) early_stop_num = 0: stop at idx 0, only the prompt comes back (t2s_model.py:747)
    r = engine.infer(ids, bert, prompt, top_k=5, early_stop_num=0, eos_suppress_steps=1)
    assert r.idx == [0, 0] and all(s.shape[0] == 1 for s in r.sequences())
    # (c) step cap: max_steps = 5 without early stop -> idx 4 (the reference's idx == 1499 rule, :747)
    r = engine.infer(ids, bert, prompt, top_k=5, early_stop_num=-1, eos_suppress_steps=11, max_steps=5)
    assert r.idx == [4, 4] and all(s.shape[0] == 1 + 4 for s in r.sequences())
    # (d) top_k > vocab, temperature 0 -> clamp 1e-5 -> the sample is the arg-max of the penalised logits
    r = engine.infer(ids, bert, prompt, top_k=5000, top_p=1.0, temperature=0.0, repetition_penalty=1.35,
                     early_stop_num=5, eos_suppress_steps=1, capture_logits=6, seed=3)
    lg = r.logits.cpu().numpy()
    for b in range(2):
        hist = list(map(int, promptc[b].numpy()))
        for s in range(r.idx[b] + 1):
            row = lg[s, b, : (1024 if s < 1 else 1025)].copy()
            seen = np.unique(np.array(hist, dtype=np.int64))
            row[seen] = np.where(row[seen] < 0, row[seen] * np.float32(1.35), row[seen] / np.float32(1.35))
            top2 = np.sort(row)[-2:]
            if top2[1] - top2[0] > 1e-3:
                assert int(r.sampled[b, s]) == int(np.argmax(row)), (b, s)
            hist.append(int(r.sampled[b, s]))
    # (e) top_p = 0 keeps only the arg-max (utils.py:172-173: first sorted position always kept)
    r = engine.infer(ids, bert, prompt, top_k=50, top_p=0.0, temperature=1.0, repetition_penalty=1.0,
                     early_stop_num=4, eos_suppress_steps=1, capture_logits=5, seed=4)
    lg = r.logits.cpu().numpy()
    for b in range(2):
        for s in range(r.idx[b] + 1):
            row = lg[s, b, : (1024 if s < 1 else 1025)]
            top2 = np.sort(row)[-2:]
            if top2[1] - top2[0] > 1e-3:
                assert int(r.sampled[b, s]) == int(np.argmax(row))  # penalty off: raw logits
    # (f) EOS forced at step 1 for utterance 0 only: idx 1, one kept token; utterance 1 continues
    forced = torch.tensor([[7, 1024, 0, 0], [8, 9, 10, 11]], dtype=torch.int32)
    r = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=3, eos_suppress_steps=1, forced=forced)
    assert r.idx == [1, 3]
    assert r.sequences()[0].cpu().tolist() == [int(promptc[0, 0]), 7]
    assert r.sequences()[1].cpu().tolist() == [int(promptc[1, 0]), 8, 9, 10]


def test_argument_validation(engine):
    """Error behaviour mirrors the reference where it has one (top_k <= 0 dies in torch.topk) and is explicit
    elsewhere; every failure is a Python exception, never a silent fallback."""
    ids, lens, prompt, bert = synthetic.make_inputs(2, [4, 6], 3, seed=41)
    ids = [t.cuda() for t in ids]
    bert = [t.cuda() for t in bert]
    prompt = prompt.cuda()
    with pytest.raises(RuntimeError, match="top_k"):
        engine.infer(ids, bert, prompt, top_k=0)
    with pytest.raises(ValueError):
        engine.infer(ids, bert[:1], prompt)
    with pytest.raises(ValueError):
        engine.infer(ids, [bert[0], bert[1][:, :5]], prompt)
    with pytest.raises(ValueError):
        engine.infer(ids, [b.cpu() for b in bert], prompt)
    with pytest.raises(ValueError):
        engine.infer([], [], None)
    with pytest.raises(RuntimeError, match="positional table"):
        engine.infer(ids, bert, prompt, top_k=1, max_steps=5000)
    with pytest.raises(RuntimeError, match="repetition_penalty"):
        engine.infer(ids, bert, prompt, top_k=1, repetition_penalty=0.0)
    # ids outside the embedding tables: the reference's nn.Embedding raises IndexError; here the device flags it (no silent
    # out-of-bounds read) and the call fails
    bad_ids = [ids[0].clone(), ids[1].clone()]
    bad_ids[1][2] = 732
    with pytest.raises(RuntimeError, match="out of range"):
        engine.infer(bad_ids, bert, prompt, top_k=1, early_stop_num=2)
    bad_ids[1][2] = -1
    with pytest.raises(RuntimeError, match="out of range"):
        engine.infer(bad_ids, bert, prompt, top_k=1, early_stop_num=2)
    bad_prompt = prompt.clone()
    bad_prompt[0, 1] = 1025
    with pytest.raises(RuntimeError, match="out of range"):
        engine.infer(ids, bert, bad_prompt, top_k=1, early_stop_num=2)
    with pytest.raises(RuntimeError, match="out of range"):
        engine.infer(ids, bert, prompt, top_k=1, early_stop_num=3, forced=torch.tensor([[1, 2, 4000], [1, 2, 3]], dtype=torch.int32))
    # early_stop_num < -1 behaves like the reference's `!= -1 and len > early_stop_num`: stop at the first step
    r = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=-5)
    assert r.idx == [0, 0]
    # the engine is still usable after errors
    r = engine.infer(ids, bert, prompt, top_k=1, early_stop_num=2)
    assert r.idx == [2, 2]


def test_codes_to_latent(engine, golden_dir):
    """t2s_codes_to_latent against the reference-recorded goldens and the oracle: bit-exact; the model's
    real shape (1024 codes x 768 channels, an utterance of 1500 tokens); lengths that are not tile multiples; the reference's
    [1, 1, T] form; an out-of-range code is an error as in the reference's embedding lookup; T = 0."""
    from oracle.t2s_oracle import codes_to_latent
    g = np.load(os.path.join(golden_dir, "latent.npz"))
    for tag in "abc":
        got = engine.codes_to_latent(torch.from_numpy(g["codes_" + tag]).cuda(), torch.from_numpy(g["codebook_" + tag]).cuda())
        assert np.array_equal(got.cpu().numpy(), g["latent_" + tag])
    rng = np.random.default_rng(5)
    cb = rng.standard_normal((1024, 768), dtype=np.float32)
    cbd = torch.from_numpy(cb).cuda()
    for T, up in ((1500, 2), (31, 2), (65, 1), (100, 3)):
        codes = rng.integers(0, 1024, T)
        got = engine.codes_to_latent(torch.from_numpy(codes).cuda(), cbd, upsample=up)
        assert got.shape == (1, 768, up * T) and np.array_equal(got.cpu().numpy(), codes_to_latent(codes, cb, up))
    assert engine.codes_to_latent(torch.zeros(0, dtype=torch.int64).cuda(), cbd).shape == (1, 768, 0)
    with pytest.raises(RuntimeError, match="outside"):
        engine.codes_to_latent(torch.tensor([3, 1024]).cuda(), cbd)  # EOS is never a code
    with pytest.raises(TypeError):
        engine.codes_to_latent(torch.tensor([3], dtype=torch.int32).cuda(), cbd)
