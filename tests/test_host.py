"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol that
include/t2s_b200.h declares (no compute calls without a GPU), error behaviour without a device, the
synthetic generator, and the multi-GPU utterance sharding over gloo (world_size 2)."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpt_sovits_b200 import _lib, shard, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "t2s_b200.h")).read()
    declared = set(re.findall(r"\b(t2s_[a-z_]+)\s*\(", header))
    declared -= {"t2s_engine"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    _lib.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    # sizes the C side compiles to (x86-64 SysV): guards the ctypes mirror against drift
    assert C.sizeof(_lib.ModelConfig) == 40
    assert C.sizeof(_lib.Request) == 120
    assert C.sizeof(_lib.Stats) == 80
    assert _lib.Request.prompt_row_stride.offset == 64 and _lib.Request.seed.offset == 104


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_device_no_fallback():
    import gpt_sovits_b200 as gsb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gsb.T2SEngine(synthetic.S1V2_CONFIG)
    lib = _lib.load()
    cfg = _lib.ModelConfig(n_layer=24, d_model=512, n_head=16, d_ff=2048, vocab=1025, phoneme_vocab=732,
                           bert_dim=1024, eos=1024, pe_len=4000, max_batch=32)
    h = C.c_void_p()
    assert lib.t2s_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.t2s_last_error() or b"CUDA" in lib.t2s_last_error()
    bad = _lib.ModelConfig(n_layer=24, d_model=1024, n_head=16, d_ff=4096, vocab=1025, phoneme_vocab=732,
                           bert_dim=1024, eos=1024, pe_len=4000, max_batch=32)
    assert lib.t2s_create(C.byref(bad), C.byref(h)) != 0
    assert b"specialised" in lib.t2s_last_error()


def test_synthetic_is_deterministic_and_bf16_exact():
    a = synthetic.make_state_dict(seed=0, config={"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=2)})
    b = synthetic.make_state_dict(seed=0, config={"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=2)})
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k])
        assert torch.equal(a[k], a[k].bfloat16().float()), k  # bf16-representable
    assert a["h.layers.1.self_attn.in_proj_weight"].shape == (1536, 512)
    # pins the generator stream the goldens were made with
    assert abs(float(a["ar_predict_layer.weight"].double().abs().sum()) - 18545.0) < 200.0
    ids, lens, prompt, bert = synthetic.make_inputs(3, [5, 9, 7], 11, seed=1)
    assert [t.shape[0] for t in ids] == [5, 9, 7] and prompt.shape == (3, 11) and prompt.stride(0) == 0
    assert bert[1].shape == (1024, 9)
    assert synthetic.config_lens(4, 60, 120, seed=2) == synthetic.config_lens(4, 60, 120, seed=2)


def test_partition_is_balanced_and_complete():
    costs = [120, 60, 61, 119, 90, 90, 75, 100, 64]
    for world in (1, 2, 4, 8):
        parts = shard.partition_utterances(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        if world <= 4:
            assert max(loads) - min(loads) <= max(costs)
    assert shard.partition_utterances([], 2) == [[], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, costs, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def fake_infer(indices):  # stands in for infer_panel_batch_infer on this rank's GPU
        calls.append(list(indices))
        ys = [torch.arange(3 + i, dtype=torch.int64) * (i + 1) for i in indices]
        return ys, [i % 5 + 1 for i in indices]

    y_list, idx_list = shard.sharded_infer(fake_infer, costs)
    out_q.put((rank, calls, [y.tolist() for y in y_list], idx_list))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_infer_gloo_world2():
    """N>1 path on CPU: two ranks, disjoint utterance shares, results identical on both ranks and in the
    original order; the only communication is the final gather."""
    costs = [80, 60, 120, 61, 95, 70, 110]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, costs, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    (r0, c0, y0, i0), (r1, c1, y1, i1) = results
    assert y0 == y1 and i0 == i1
    assert len(c0) == 1 and len(c1) == 1 and sorted(c0[0] + c1[0]) == list(range(7)) and not set(c0[0]) & set(c1[0])
    for i in range(7):
        assert y0[i] == (np.arange(3 + i) * (i + 1)).tolist() and i0[i] == i % 5 + 1


def _write_ckpt(path, n_layer=2, drop=None, model_overrides=None):
    cfg = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=n_layer, linear_units=2048, random_bert=0), "data": {"max_sec": 54}}
    if model_overrides:
        cfg["model"].update(model_overrides)
    sd = synthetic.make_state_dict(seed=3, config={"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=n_layer)})
    weight = {"model." + k: v.half() for k, v in sd.items() if k != drop}  # fp16 + Lightning prefix, as process_ckpt.py writes
    torch.save({"weight": weight, "config": cfg, "info": "GPT-e1"}, path)
    return sd


def test_checkpoint_reader(tmp_path):
    """The s1 checkpoint wire format of the reference (process_ckpt.py:12-17, read by TTS.py:585-599): prefix stripped,
    shapes validated, unsupported architectures and incomplete files rejected with a clear message."""
    import gpt_sovits_b200 as gsb
    p = str(tmp_path / "s1.ckpt")
    sd = _write_ckpt(p)
    config, got = gsb.read_checkpoint(p)
    assert config["data"]["max_sec"] == 54 and config["model"]["n_layer"] == 2
    assert set(got) >= set(sd)
    assert got["h.layers.1.linear1.weight"].dtype == torch.float16
    torch.testing.assert_close(got["ar_predict_layer.weight"].float(), sd["ar_predict_layer.weight"].half().float())
    _write_ckpt(p, drop="h.layers.1.norm2.bias")
    with pytest.raises(ValueError, match="norm2.bias"):
        gsb.read_checkpoint(p)
    _write_ckpt(p, model_overrides={"hidden_dim": 1024, "embedding_dim": 1024})
    with pytest.raises(ValueError, match="hidden_dim"):
        gsb.read_checkpoint(p)
    torch.save({"something": 1}, p)
    with pytest.raises(ValueError, match="not an s1 checkpoint"):
        gsb.read_checkpoint(p)


def test_patch_real_reference_class():
    """Boundary check against the REAL Text2SemanticDecoder (needs /root/reference, i.e. the build container): the class-level
    patch installs four methods whose signatures match the reference's, TTS.run's per-request instance rebinding
    (TTS.py:1042-1047) resolves to the patch, and on a CPU model the patched methods raise (no CPU fallback)."""
    import inspect

    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    import gpt_sovits_b200 as gsb
    from gpt_sovits_b200 import decoder
    ref = ref_harness.import_reference()
    cls = ref.Text2SemanticDecoder
    names = ["infer_panel", "infer_panel_naive", "infer_panel_naive_batched", "infer_panel_batch_infer"]
    orig = {n: cls.__dict__[n] for n in names}
    sigs = {n: inspect.signature(orig[n]) for n in names}
    gsb.patch_reference(cls)
    try:
        for n in names:
            assert cls.__dict__[n] is getattr(decoder, n)
            got = inspect.signature(cls.__dict__[n])
            assert list(got.parameters) == list(sigs[n].parameters), n
            for k, p in sigs[n].parameters.items():
                assert got.parameters[k].default == p.default, (n, k)
                assert got.parameters[k].kind == p.kind, (n, k)
        cfg = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=1)}
        model = cls(cfg).eval()
        # what TTS.run does at every request: rebinding on the INSTANCE picks up the class-level patch
        model.infer_panel = model.infer_panel_batch_infer
        assert model.infer_panel.__func__ is decoder.infer_panel_batch_infer
        model.infer_panel = model.infer_panel_naive_batched
        assert model.infer_panel.__func__ is decoder.infer_panel_naive_batched
        ids, lens, prompt, bert = synthetic.make_inputs(2, [4, 6], 3, seed=1)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.infer_panel(ids, lens, prompt, bert, top_k=5, top_p=1, temperature=1.0, early_stop_num=4)
        with pytest.raises(RuntimeError, match="top_k"):
            model.infer_panel_naive(ids[0][None], lens[:1], prompt[:1], bert[0][None], top_k=0)
        with pytest.raises(RuntimeError, match="batch size must be 1"):  # checked before any engine work
            model.infer_panel_naive(torch.stack([ids[0], ids[0]]), lens[:1], prompt, torch.stack([bert[0], bert[0]]), top_k=5)
    finally:
        gsb.unpatch_reference(cls)
    for n in names:
        assert cls.__dict__[n] is orig[n]


def test_build_dependencies_cover_every_kernel_header():
    """_lib.build() must rebuild after an edit to ANY header engine.cu includes (cluster_decode.cuh was missing)."""
    src = open(os.path.join(ROOT, "gpt-sovits_b200", "csrc", "engine.cu")).read()
    deps = {os.path.basename(p) for p in _lib.SOURCES}
    for inc in re.findall(r'#include "([^"]+)"', src):
        assert os.path.basename(inc) in deps, inc
    for h in os.listdir(os.path.join(ROOT, "gpt-sovits_b200", "csrc")):
        if h.endswith(".cuh"):
            assert h in deps, h


def test_oracle_session_slices_equal_one_shot(pe_table):
    """The resumable form of the oracle loop (used by bench.py's CPU arms and the streaming tests): run(n) slices give exactly
    what one generate() call gives."""
    from oracle.t2s_oracle import OracleSession, T2SOracle
    cfg = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=2)}
    sd = synthetic.make_state_dict(seed=3, config=cfg, eos_scale=2.5)
    o = T2SOracle(sd, pe_table)
    ids, lens, prompt, bert = synthetic.make_inputs(3, [5, 9, 7], 6, seed=8)
    args = ([t.numpy() for t in ids], [t.numpy() for t in bert], prompt.numpy())
    kw = dict(top_k=5, early_stop_num=25, eos_window=1, seed=11)
    whole = o.generate(*args, **kw)
    s = OracleSession(o, *args, **kw)
    ran = 0
    while s.active:
        mid = s.result()
        assert all((i == -1) == (b in s.active) for b, i in enumerate(mid["idx"]))
        ran += s.run(4)
    part = s.result()
    assert part["idx"] == whole["idx"] and ran == max(whole["idx"]) + 1
    for a, b in zip(part["tokens"], whole["tokens"]):
        np.testing.assert_array_equal(a, b)


def test_bench_cpu_arm_port_is_measured_not_projected(monkeypatch):
    """bench.py --impl reference on a box without the reference tree (the GPU box): the port arm times real decode steps of a
    resident session; tokens per sample = batch x n, nothing extrapolated."""
    import bench
    monkeypatch.setenv("T2S_BENCH_FORCE_PORT", "1")
    w = dict(batch=2, lo=6, hi=9, prompt=8, top_k=15, top_p=1.0, temperature=1.0, repetition_penalty=1.35, cap=40, eos_window=1)
    monkeypatch.setattr(bench.synthetic, "make_state_dict", _small_sd)  # 2 layers keep the CPU test quick
    arm = bench.CpuArm(w, "tiny")
    assert arm.kind == "port" and arm.cores == (os.cpu_count() or 1)
    n = arm.calibrate(4, 8.0)
    assert 1 <= n <= (40 - 8) // 4
    toks, secs = arm.sample()
    assert toks == 2 * n and secs > 0
    toks2, _ = arm.sample()
    assert toks2 == 2 * n
    assert "measured decode steps" in arm.describe() and "prefill" in arm.describe()


def _small_sd(**kw):
    kw = dict(kw)
    kw["config"] = {"model": dict(synthetic.S1V2_CONFIG["model"], n_layer=2)}
    return _ORIG_MAKE_SD(**kw)


_ORIG_MAKE_SD = synthetic.make_state_dict


def test_bucket_batches_matches_reference_to_batch(golden_dir):
    """bucket_batches / recovery_order against batch index lists recorded from the reference's own TTS.to_batch
    (oracle/make_goldens.py case_to_batch): same batches, same order inside each batch, ties and thresholds included."""
    import json

    import gpt_sovits_b200 as gsb
    cases = json.load(open(os.path.join(golden_dir, "to_batch.json")))
    assert len(cases) >= 10
    for c in cases:
        got = gsb.bucket_batches(c["lengths"], c["batch_size"], c["threshold"], c["split_bucket"])
        assert got == c["batch_index_list"], c
        assert sorted(i for b in got for i in b) == list(range(len(c["lengths"])))
        back = gsb.recovery_order([[c["lengths"][i] for i in b] for b in got], got)
        assert back == c["lengths"]
    assert gsb.bucket_batches([], 4) == []
    with pytest.raises(ValueError):
        gsb.bucket_batches([1, 2], 0)
    with pytest.raises(ValueError):
        gsb.recovery_order([[1]], [[0, 1]])


def test_streaming_session_admission_policy():
    """StreamingSession's host logic with a fake engine (no GPU): every utterance is served exactly once, at most `slots` decode
    at the same time, released slots are reused, and with admit_min = n waiting utterances are admitted in groups of >= n (or all
    that wait) instead of one prefill per freed slot."""
    import gpt_sovits_b200 as gsb

    class FakeResult:
        def __init__(self, toks, idx):
            self._t, self.idx, self.logits = toks, idx, None

        def sequences(self):
            return self._t

    class FakeEngine:
        """utterance k needs 3 + 2 * (k % 5) steps; tokens = [k] * steps"""
        def __init__(self):
            self.slot_utt, self.slot_left, self.free, self.admissions, self.max_active = {}, {}, [], [], 0

        def _place(self, keys, slots):
            for k, s in zip(keys, slots):
                self.slot_utt[s], self.slot_left[s] = k, 3 + 2 * (k % 5)
            self.admissions.append(len(keys))
            self.max_active = max(self.max_active, sum(1 for v in self.slot_left.values() if v > 0))

        def infer(self, ids, bert, prompt, max_new_steps=0, reserve_slots=0, reserve_positions=0, utt_ids=None, **kw):
            self.cap = reserve_slots
            self._place(list(utt_ids), list(range(len(ids))))
            return FakeResult(None, None)

        def admit(self, ids, bert, prompt, utt_ids=None):
            slots = []
            for _ in ids:
                slots.append(self.free.pop(0) if self.free else len(self.slot_utt))
                assert slots[-1] < self.cap
            self._place(list(utt_ids), slots)
            return slots

        def session_result(self):
            n = max(self.slot_utt) + 1
            toks = [torch.full((3 + 2 * (self.slot_utt.get(s, 0) % 5),), self.slot_utt.get(s, 0), dtype=torch.int64) for s in range(n)]
            return FakeResult(toks, [(3 + 2 * (self.slot_utt[s] % 5)) if s in self.slot_utt and self.slot_left[s] <= 0 else -1 for s in range(n)])

        def release(self, slots):
            for s in slots:
                self.slot_left[s] = 10 ** 9  # empty until re-admitted
            self.free = sorted(set(self.free) | set(slots))

        def decode_more(self, n):
            for s in self.slot_left:
                if self.slot_left[s] < 10 ** 8:
                    self.slot_left[s] -= n

    for admit_min in (1, 4):
        eng = FakeEngine()
        sess = gsb.StreamingSession(eng, slots=6, slice_steps=2, admit_min=admit_min, top_k=1)
        n = 23
        keys = sess.submit([torch.zeros(4, dtype=torch.int64)] * n, [torch.zeros(1024, 4)] * n, torch.zeros(n, 5, dtype=torch.int64))
        assert keys == list(range(n))
        got = {}
        for key, toks, idx in sess:
            assert key not in got
            got[key] = (toks, idx)
        assert sorted(got) == list(range(n))
        for k, (toks, idx) in got.items():
            assert idx == 3 + 2 * (k % 5) and bool((toks == k).all())
        assert eng.max_active <= 6 and sum(eng.admissions) == n
        if admit_min == 4:  # groups of >= 4 except the tail of the queue
            assert all(a >= 4 for a in eng.admissions[:-1]), eng.admissions
            assert len(eng.admissions) < n - 6
