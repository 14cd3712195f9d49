"""CPU: the numpy oracle (oracle/) against the golden vectors recorded from the live reference
(tests/golden/*.npz, made by oracle/make_goldens.py).  Tolerance: the reference's own fp32 noise
floor against itself is 4-5e-5 on logits; 2e-4 is allowed here because
numpy/OpenBLAS and ATen sum in different orders."""
import os

import numpy as np
import pytest

from gpt_sovits_b200 import synthetic
from oracle import sampler_oracle as so
from oracle.t2s_oracle import EOS_WINDOW_BATCH, EOS_WINDOW_NAIVE, T2SOracle

TOL = 2e-4


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _inputs(g):
    L = [int(v) for v in g["phoneme_lens"]]
    ids, lens, prompt, bert = synthetic.make_inputs(len(L), L, int(g["prompt_len"]), seed=int(g["input_seed"]))
    return ([t.numpy() for t in ids], [t.numpy() for t in bert],
            None if prompt is None else prompt.numpy())


def _check_logits(out, g, eos_window):
    ref = g["logits"]
    assert len(out["logits"]) == ref.shape[0]
    worst = 0.0
    for s, lg in enumerate(out["logits"]):
        width = 1024 if s < eos_window else 1025
        r = ref[s, : lg.shape[0], :width]
        assert not np.isnan(r).any()
        worst = max(worst, float(np.abs(lg[:, :width] - r).max()))
    assert worst < TOL, worst


def test_naive_b1(golden_dir, weights_seed0, pe_table):
    g = _load(golden_dir, "naive_b1")
    ids, bert, prompt = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
                     early_stop_num=int(g["early_stop_num"]), eos_window=EOS_WINDOW_NAIVE, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_NAIVE)
    assert out["idx"] == [int(g["idx"])]
    np.testing.assert_array_equal(out["tokens"][0], g["y"][0])


def test_batch_b4(golden_dir, weights_seed0, pe_table):
    g = _load(golden_dir, "batch_b4")
    ids, bert, prompt = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]),
                     eos_window=EOS_WINDOW_BATCH, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_BATCH)
    assert out["idx"] == [int(v) for v in g["idx"]]
    for b in range(4):
        np.testing.assert_array_equal(out["tokens"][b], g["y"][b])


def test_retire_b6(golden_dir, pe_table):
    g = _load(golden_dir, "retire_b6")
    ids, bert, prompt = _inputs(g)
    sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), eos_scale=float(g["eos_scale"]))
    o = T2SOracle(sd, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]),
                     eos_window=EOS_WINDOW_BATCH, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_BATCH)
    assert out["idx"] == [int(v) for v in g["idx"]]
    assert len(set(out["idx"])) > 2  # retirement at different steps is actually exercised
    for b in range(6):
        y = g["y"][b]
        np.testing.assert_array_equal(out["tokens"][b], y[y >= 0])


def test_reffree_b1(golden_dir, weights_seed0, pe_table):
    g = _load(golden_dir, "reffree_b1")
    ids, bert, _ = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    y, idx = o.infer_panel_naive(ids[0][None], None, None, bert[0][None], top_k=1, top_p=1.0,
                                 early_stop_num=int(g["early_stop_num"]), temperature=1.0,
                                 repetition_penalty=1.35)
    assert idx == 0 == int(g["idx"])
    np.testing.assert_array_equal(y, g["y"])


# ---- round 2: the real horizons (long contexts, config 2's batch, the naive-batched loop, fp16 checkpoints) ----------------
def test_long_b2(golden_dir, weights_seed0, pe_table):
    """S0 = 900 (300 phonemes + 600 prompt tokens, the config-5 shape): 8 K/V pages per sequence."""
    g = _load(golden_dir, "long_b2")
    ids, bert, prompt = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]),
                     eos_window=EOS_WINDOW_BATCH, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_BATCH)
    assert out["idx"] == [int(v) for v in g["idx"]]
    for b in range(2):
        np.testing.assert_array_equal(out["tokens"][b], g["y"][b])


def test_long_b1(golden_dir, weights_seed0, pe_table):
    """S0 = 1470 -> 1510: the KV length crosses 1500 positions (12 pages, partial last page)."""
    g = _load(golden_dir, "long_b1")
    ids, bert, prompt = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]),
                     eos_window=EOS_WINDOW_NAIVE, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_NAIVE)
    assert out["idx"] == [int(g["idx"])]
    np.testing.assert_array_equal(out["tokens"][0], g["y"][0])


@pytest.mark.parametrize("name", ["naive_batched_b4", "naive_batched_reffree"])
def test_naive_batched(golden_dir, pe_table, name):
    """infer_panel_naive_batched (t2s_model.py:781-812): multi-item, 11-step EOS window per item, with and without prompts."""
    g = _load(golden_dir, name)
    ids, bert, prompt = _inputs(g)
    sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), eos_scale=float(g["eos_scale"]))
    o = T2SOracle(sd, pe_table)
    ys, idxs = o.infer_panel_naive_batched(ids, None, prompt, bert, top_k=1, top_p=1.0,
                                           early_stop_num=int(g["early_stop_num"]), temperature=1.0, repetition_penalty=1.35)
    assert idxs == [int(v) for v in g["idx"]]
    for b, y in enumerate(ys):
        ref = g["y"][b]
        np.testing.assert_array_equal(y, ref[ref >= 0])
    # per-step logits of every item, through the one ragged batch the CUDA path runs (same semantics: sequences never interact)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_window=EOS_WINDOW_NAIVE,
                     record_logits=True)
    worst = 0.0
    for s, (lg, act) in enumerate(zip(out["logits"], out["active"])):
        w = 1024 if s < EOS_WINDOW_NAIVE else 1025
        for r, b in enumerate(act):
            assert s < int(g["n_steps"][b])
            worst = max(worst, float(np.abs(lg[r, :w] - g["logits"][s, b, :w]).max()))
    assert worst < TOL, worst
    if prompt is not None:
        assert len(set(idxs)) > 2  # the items stop at different steps


def test_cfg2_b32(golden_dir, weights_seed0, pe_table):
    """BASELINE config 2 (B=32, 60..120 phonemes + 150 prompt, top_k=15 sampled by the reference): 16 steps, teacher-forced
    with the tokens the reference emitted."""
    g = _load(golden_dir, "cfg2_b32")
    ids, bert, prompt = _inputs(g)
    o = T2SOracle(weights_seed0, pe_table)
    out = o.generate(ids, bert, prompt, top_k=15, early_stop_num=int(g["early_stop_num"]), eos_window=EOS_WINDOW_BATCH,
                     forced=g["emitted"], record_logits=True)
    _check_logits(out, g, EOS_WINDOW_BATCH)
    assert out["idx"] == [int(v) for v in g["idx"]]
    for b in range(32):
        np.testing.assert_array_equal(out["tokens"][b], g["y"][b])


def test_fp16_weights(golden_dir, pe_table):
    g = _load(golden_dir, "fp16w_b1")
    ids, bert, prompt = _inputs(g)
    sd = synthetic.make_state_dict(seed=0, rounding="fp16")
    assert any(not np.array_equal(v.numpy(), synthetic.bf16_round(v).numpy()) for v in sd.values())  # NOT bf16-representable
    o = T2SOracle(sd, pe_table)
    out = o.generate(ids, bert, prompt, top_k=1, early_stop_num=int(g["early_stop_num"]), eos_window=EOS_WINDOW_NAIVE,
                     record_logits=True)
    _check_logits(out, g, EOS_WINDOW_NAIVE)
    np.testing.assert_array_equal(out["tokens"][0], g["y"][0])


def test_ckpt_s1_v1_architecture(golden_dir, pe_table):
    """configs/s1.yaml (12 layers, 512 phonemes), fp16 checkpoint values: the other 512-d member of the s1 family."""
    g = _load(golden_dir, "ckpt_s1v1")
    cfg = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=int(g["n_layer"]), phoneme_vocab_size=int(g["phoneme_vocab"]))}
    sd = synthetic.make_state_dict(seed=int(g["weight_seed"]), config=cfg, rounding="fp16")
    L = [int(v) for v in g["phoneme_lens"]]
    ids, lens, prompt, bert = synthetic.make_inputs(len(L), L, int(g["prompt_len"]), seed=int(g["input_seed"]), phoneme_vocab=512)
    o = T2SOracle(sd, pe_table)
    assert o.L == 12
    out = o.generate([t.numpy() for t in ids], [t.numpy() for t in bert], prompt.numpy(), top_k=1,
                     early_stop_num=int(g["early_stop_num"]), eos_window=EOS_WINDOW_BATCH, record_logits=True)
    _check_logits(out, g, EOS_WINDOW_BATCH)
    assert out["idx"] == [int(v) for v in g["idx"]]
    for b in range(len(L)):
        np.testing.assert_array_equal(out["tokens"][b], g["y"][b])


def test_sampler_kat(golden_dir):
    g = _load(golden_dir, "sampler_kat")
    for i in range(g["logits"].shape[0]):
        w = int(g["width"][i])
        temperature, top_k, top_p, rp = [float(v) for v in g["params"][i]]
        prev = g["prev"][i]
        prev = prev[prev >= 0]
        lg = g["logits"][i, :w].copy()
        probs = so.logits_to_probs(lg, prev, temperature, int(top_k), top_p, rp)
        np.testing.assert_allclose(lg, g["penalised"][i, :w], rtol=0, atol=0)  # in-place penalty, exact
        ref = g["probs"][i, :w]
        assert ((probs > 0) == (ref > 0)).all(), f"row {i}: support differs"
        np.testing.assert_allclose(probs, ref, rtol=2e-6, atol=1e-9)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    z = np.zeros(1, np.uint32)
    r = so.philox4x32_10(z, z, z, z, 0, 0)
    assert [int(v[0]) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.full(1, 0xFFFFFFFF, np.uint32)
    r = so.philox4x32_10(f, f, f, f, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v[0]) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = so.philox4x32_10(np.array([0x243F6A88], np.uint32), np.array([0x85A308D3], np.uint32),
                         np.array([0x13198A2E], np.uint32), np.array([0x03707344], np.uint32),
                         0xA4093822, 0x299F31D0)
    assert [int(v[0]) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_exp_noise_properties():
    q = so.exp_noise(1234, 3, 17, 1025)
    assert q.dtype == np.float32 and q.shape == (1025,) and (q > 0).all() and np.isfinite(q).all()
    assert abs(float(so.exp_noise(1, 0, 0, 200000).mean()) - 1.0) < 0.02
    assert not np.array_equal(q, so.exp_noise(1234, 3, 18, 1025))
    assert not np.array_equal(q, so.exp_noise(1234, 4, 17, 1025))


def test_codes_to_latent_golden(golden_dir):
    """The restated first op after the path (quantizer.decode + nearest x2, module/models.py:989-991) against outputs of the
    reference's own ResidualVectorQuantizer + F.interpolate (oracle/make_goldens.py case_latent): bit-exact (a gather)."""
    from oracle.t2s_oracle import codes_to_latent
    g = np.load(os.path.join(golden_dir, "latent.npz"))
    for tag in "abc":
        got = codes_to_latent(g["codes_" + tag], g["codebook_" + tag], 2)
        assert got.dtype == np.float32 and np.array_equal(got, g["latent_" + tag])
    with pytest.raises(IndexError):
        codes_to_latent(np.array([0, 40]), g["codebook_a"], 2)
    assert codes_to_latent(np.zeros(0, np.int64), g["codebook_a"], 2).shape == (1, 48, 0)
