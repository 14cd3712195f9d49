#!/usr/bin/env python
"""bench.py — semantic tokens/s of the T2S decode hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

Workload (config.workload), default cfg2_b32 = BASELINE.json configs[1]: s1-v2 (24L/512d/16h/FFN2048, random-init
"sensitive" synthetic weights, EOS row zeroed so every sequence runs to the cap), batch 32 utterances per GPU
(60..120 phonemes + 150 prompt tokens), top_k=15 top_p=1.0 temperature=1.0 repetition_penalty=1.35, 1000-step cap.
One bench "step" = one infer_panel_batch_infer call over that batch: prefill + decode + sampling + result
(32,000 semantic tokens per GPU).

  value     whole-job kept tokens / device time, inputs already resident in HBM (CUDA events on the launching stream,
            barrier + synchronize on both sides, max over ranks)
  e2e       the same metric through the public API with HOST buffers (pinned): H2D of phoneme ids, BERT features and
            prompt, and D2H of the token matrix + idx, inside the timed region
  roofline  the persistent decode kernel: algorithmic bytes (bf16 weights once per step + every active sequence's KV
            read + its new KV row) / its CUDA-event duration, against the measured HBM peak
  sweep     (N = 1) short runs of the other BASELINE configs on the same GPU - cfg1_b1, batch 8 / 64 / 256 (config 4),
            cfg5_b128 - each with the decode launch time, its algorithmic bytes and roofline fraction, and prefill time
  cfg3      BASELINE config 3 as a STRONG-scaling job: 120 ragged utterances (fixed total) with natural EOS, sharded over
            the N ranks by gpt_sovits_b200.shard.sharded_infer (LPT by phoneme count, one all_gather_object at the end)
  cpu_baseline  the reference path on the host cores, on a bounded sample (see cpu_arm)

--impl reference: the reference's own CPU implementation of the path on the box's host cores, all threads, same
workload / metric / unit.  The unmodified reference (Text2SemanticDecoder.infer_panel_batch_infer, imported from
/root/reference or baseline/_ref with a torchmetrics stub) is used when one of those trees exists at run time
(kind "reference"); the GPU box has neither, so there the numpy oracle port runs (kind "port").  Every step is a
MEASURED bounded sample - nothing is projected: the port keeps ONE resident session (prefill during warm-up) and
each step decodes `n` further tokens for all utterances (n calibrated in the first warm-up step so that the
whole run fits ~2 minutes); the real reference has no resumable loop, so each of its steps is one whole call with
early_stop_num = n (prefill included, which is why its tokens/s is lower on short samples: stated in `sample`).

Multi-GPU: utterances are independent, so ranks get disjoint batches (weak scaling), no collective on the data
path; NCCL is used only for the barrier and the max-over-ranks of the timing (and cfg3's final gather).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gpt_sovits_b200 as gsb  # noqa: E402
from gpt_sovits_b200 import synthetic  # noqa: E402

METRIC = "semantic_tokens_per_sec"
UNIT = "tokens/s"
TOKENS_PER_SEC_AUDIO = 25.0

WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2_b32": dict(batch=32, lo=60, hi=120, prompt=150, top_k=15, top_p=1.0, temperature=1.0,
                     repetition_penalty=1.35, cap=1000, eos_window=1),
    # BASELINE.json configs[0] (the reference's CPU-runnable case): B=1, greedy, 500 steps, naive path
    "cfg1_b1": dict(batch=1, lo=80, hi=80, prompt=150, top_k=1, top_p=1.0, temperature=1.0,
                    repetition_penalty=1.35, cap=500, eos_window=11),
    # BASELINE.json configs[3]: batch sweep per GPU, config-2 sampling, 500-step cap
    "cfg4_b8": dict(batch=8, lo=60, hi=120, prompt=150, top_k=15, top_p=1.0, temperature=1.0,
                    repetition_penalty=1.35, cap=500, eos_window=1),
    "cfg4_b64": dict(batch=64, lo=60, hi=120, prompt=150, top_k=15, top_p=1.0, temperature=1.0,
                     repetition_penalty=1.35, cap=500, eos_window=1),
    "cfg4_b256": dict(batch=256, lo=60, hi=120, prompt=150, top_k=15, top_p=1.0, temperature=1.0,
                      repetition_penalty=1.35, cap=500, eos_window=1),
    # BASELINE.json configs[4]: long prompt stress
    "cfg5_b128": dict(batch=128, lo=300, hi=300, prompt=600, top_k=15, top_p=1.0, temperature=1.0,
                      repetition_penalty=1.35, cap=600, eos_window=1),
}
SWEEP = ["cfg1_b1", "cfg4_b8", "cfg4_b64", "cfg4_b256", "cfg5_b128"]
# BASELINE.json configs[2]: ~120 sentences of a long-form text, ragged, natural EOS (EOS row of the head scaled so that
# sequences stop on their own), fixed TOTAL work: strong scaling over the ranks
# (greedy, as SURVEY.md 8d prescribes for this config: the tokens - and with them the total work - do not depend on how the
#  utterances are sharded; measured on B200 with this head: 122 tokens per utterance on average, 2..426)
CFG3 = dict(total=120, lo=60, hi=140, prompt=150, top_k=1, top_p=1.0, temperature=1.0, repetition_penalty=1.35,
            cap=1000, eos_window=1, eos_scale=1.4, weight_seed=3,
            slots=56, slice_steps=24, admit_min=8)  # the continuous-batching schedule (StreamingSession)


def workload_inputs(w, rank):
    L = synthetic.config_lens(w["batch"], w["lo"], w["hi"], seed=100 + rank)
    ids, lens, prompt, bert = synthetic.make_inputs(w["batch"], L, w["prompt"], seed=200 + rank)
    return L, ids, lens, prompt, bert


def config_dict(w, name, n_gpus, global_batch=None):
    return {
        "workload": f"{name}: t2s s1-v2 24L/512d/16h/FFN2048 random-init (EOS row zeroed), batch {w['batch']}/GPU, "
                    f"{w['lo']}..{w['hi']} phonemes + {w['prompt']} prompt tokens, top_k={w['top_k']} top_p={w['top_p']} "
                    f"T={w['temperature']} rp={w['repetition_penalty']}, {w['cap']}-step cap; one step = one "
                    f"infer_panel_batch_infer call (prefill+decode+sampling)",
        "global_batch": w["batch"] * n_gpus if global_batch is None else global_batch,
        "steps_cap": w["cap"],
        "parallelism": f"utterance-sharded x{n_gpus}, no data-path collective",
        "l2": "no L2 flush needed: per-step working set (152 MB weights + KV of all sequences) exceeds the 126 MB L2",
    }


# ---- clocks -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        return out


# ---- CPU arms (the reference path on the host cores) ------------------------------------------------------------------
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms use every host core."""
    cores = os.cpu_count() or 1
    try:
        torch.set_num_threads(cores)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def find_reference_tree():
    """The unmodified reference, if it exists at run time: /root/reference (build container) or baseline/_ref."""
    if os.environ.get("T2S_BENCH_FORCE_PORT"):  # lets the build container exercise the GPU box's code path
        return None
    for root in ("/root/reference/GPT_SoVITS", os.path.join(ROOT, "baseline", "_ref", "GPT_SoVITS"),
                 os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(root, "AR", "models", "t2s_model.py")):
            return root
    return None


class CpuArm:
    """Bounded, MEASURED samples of a workload on the host cores.  kind = "reference": each sample is one call of the
    unmodified infer_panel_batch_infer / infer_panel with early_stop_num = n (prefill included).  kind = "port": one
    resident numpy-oracle session, each sample = n further decode steps of every utterance."""

    def __init__(self, w, name, rank=0):
        self.w, self.name = w, name
        self.cores = _use_all_host_threads()
        self.tree = find_reference_tree()
        self.kind = "reference" if self.tree else "port"
        self.L, self.ids, self.lens, self.prompt, self.bert = workload_inputs(w, rank)
        self.sd = synthetic.make_state_dict(seed=0, eos_scale=0.0)
        self.n = None
        self.session = None
        self.t_prefill = None
        if self.kind == "reference":
            from oracle import ref_harness
            ref_harness.REFERENCE_ROOT = self.tree
            self.model = ref_harness.build_reference_model(self.sd, synthetic.S1V2_CONFIG)
            self.quiet = ref_harness.quiet
        else:
            from oracle.t2s_oracle import T2SOracle
            self.oracle = T2SOracle(self.sd, synthetic.sine_pe())

    def _ref_call(self, n):
        w = self.w
        kw = dict(top_k=w["top_k"], top_p=w["top_p"], temperature=w["temperature"], early_stop_num=n,
                  repetition_penalty=w["repetition_penalty"])
        with self.quiet(), torch.no_grad():
            if w["eos_window"] == 11:  # the naive single-utterance path (config 1)
                y, idx = self.model.infer_panel(self.ids[0][None], self.lens, self.prompt, self.bert[0][None], **kw)
                return int(idx)
            ys, idxs = self.model.infer_panel_batch_infer(self.ids, self.lens, self.prompt, self.bert, max_len=max(self.L), **kw)
            return int(sum(idxs))

    def calibrate(self, n_samples, budget_s):
        """First (untimed) contact: the port prefills its session; both kinds measure a decode step and fix n."""
        w = self.w
        t0 = time.perf_counter()
        if self.kind == "port":
            from oracle.t2s_oracle import OracleSession
            self.session = OracleSession(self.oracle, [t.numpy() for t in self.ids], [t.numpy() for t in self.bert],
                                         self.prompt.numpy(), top_k=w["top_k"], top_p=w["top_p"], temperature=w["temperature"],
                                         repetition_penalty=w["repetition_penalty"], early_stop_num=w["cap"],
                                         eos_window=w["eos_window"], seed=1)
            self.t_prefill = time.perf_counter() - t0
            t1 = time.perf_counter()
            self.session.run(3)
            t_step = (time.perf_counter() - t1) / 3
            per_sample = max(budget_s - self.t_prefill, 10.0) / max(n_samples, 1)
            # numpy's decode step slows down as the K/V length grows over the session (~1.6x by the end): size for that
            self.n = int(max(2, min(64, per_sample / (1.6 * t_step))))
            self.n = max(1, min(self.n, (w["cap"] - 8) // max(n_samples, 1)))  # the session must outlast every sample
        else:
            self._ref_call(2)
            t_a = time.perf_counter() - t0
            t1 = time.perf_counter()
            self._ref_call(8)
            t_b = time.perf_counter() - t1
            t_step = max((t_b - t_a) / 6, 1e-4)
            self.t_prefill = max(t_a - 2 * t_step, 0.0)
            per_sample = budget_s / max(n_samples, 1)
            self.n = int(max(4, min(w["cap"], (per_sample - self.t_prefill) / t_step)))
        return self.n

    def sample(self):
        """One measured sample: (tokens, seconds)."""
        t0 = time.perf_counter()
        if self.kind == "port":
            before = sum(len(g) for g in self.session.gen)
            self.session.run(self.n)
            toks = sum(len(g) for g in self.session.gen) - before
            if not self.session.active or self.session.step + self.n >= self.w["cap"]:
                self.session = None  # (never reached within the time budget; a fresh session would be needed)
        else:
            toks = self._ref_call(self.n)
        return toks, time.perf_counter() - t0

    def describe(self):
        w = self.w
        if self.kind == "port":
            return (f"numpy fp32 port of the reference path (oracle/), {self.name} utterances (batch {w['batch']}), one resident "
                    f"session: prefill {self.t_prefill:.1f} s (untimed, warm-up), each step = {self.n} measured decode steps of all "
                    f"{w['batch']} utterances at KV ~{w['prompt'] + (w['lo'] + w['hi']) // 2}+ positions (the full workload's mean KV is "
                    f"longer, so this favours the CPU); BLAS threads = host cores; /root/reference absent on this box")
        return (f"UNMODIFIED reference ({self.tree}) Text2SemanticDecoder."
                f"{'infer_panel' if w['eos_window'] == 11 else 'infer_panel_batch_infer'} on CPU fp32, torch threads = host cores, "
                f"{self.name} utterances (batch {w['batch']}); each step = one whole call with early_stop_num={self.n} "
                f"(prefill ~{self.t_prefill:.1f} s included in every step, so tokens/s is lower than on the full "
                f"{w['cap']}-step workload where the prefill is amortised)")


def run_reference(args, rank, world):
    if rank != 0:
        return  # the CPU's throughput does not depend on the number of GPUs: rank 0 alone runs and prints
    name = args.workload
    w = WORKLOADS[name]
    arm = CpuArm(w, name)
    n_samples = args.warmup + args.steps
    arm.calibrate(n_samples, args.cpu_budget)
    for _ in range(args.warmup):
        arm.sample()
    toks, secs = 0, 0.0
    for _ in range(args.steps):
        t, s = arm.sample()
        toks += t
        secs += s
    value = toks / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the work this arm actually did: ONE rank's batch, whatever N is (tokens/s of the host cores is the comparison)
        "config": config_dict(w, name, args.gpus, global_batch=w["batch"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tokens_per_step": toks / args.steps,
        "rtf": TOKENS_PER_SEC_AUDIO / (value / w["batch"]),
    }
    print(json.dumps(line))


# ---- our arm ------------------------------------------------------------------------------------------------
def algorithmic_bytes(st):
    return (st["decode_steps"] * st["weight_bytes_per_step"]
            + st["kv_bytes_per_position"] * (st["decode_kv_positions"] + st["decode_tokens"]))


def sweep_point(eng, name, dev, peak, reps=2):
    """One short device-resident run of another BASELINE config: decode launch time, algorithmic bytes, roofline fraction."""
    w = WORKLOADS[name]
    L, ids, lens, prompt, bert = workload_inputs(w, 0)
    ids = [t.to(dev) for t in ids]
    bert = [t.to(dev) for t in bert]
    prompt = prompt.to(dev)
    kw = dict(top_k=w["top_k"], top_p=w["top_p"], temperature=w["temperature"], repetition_penalty=w["repetition_penalty"],
              early_stop_num=w["cap"], eos_suppress_steps=w["eos_window"], max_steps=1500)
    eng.infer(ids, bert, prompt, seed=1, **kw)  # warm-up (allocations, module load)
    dec_ms, pre_ms, by, toks, steps, kvpos, seqsteps = 0.0, 0.0, 0.0, 0, 0, 0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        r = eng.infer(ids, bert, prompt, seed=10 + i, **kw)
        st = r.stats
        dec_ms += st["decode_ms"]; pre_ms += st["prefill_ms"]; by += algorithmic_bytes(st)
        toks += sum(max(v, 0) for v in r.idx); steps += int(st["decode_steps"])
        kvpos += int(st["decode_kv_positions"]); seqsteps += int(st["decode_tokens"])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    achieved = by / (dec_ms / 1000.0) / 1e9
    return {
        "workload": name, "batch": w["batch"], "steps_cap": w["cap"], "decode_mode": int(r.stats["decode_mode"]),
        "tokens_per_s": toks / (ms / 1000.0), "ms_per_call": ms / reps, "launch_ms": dec_ms / reps,
        "prefill_ms": pre_ms / reps, "prefill_rows": int(r.stats["prefill_rows"]),
        "us_per_decode_step": 1000.0 * dec_ms / max(steps, 1), "decode_steps_per_call": steps / reps,
        "mean_kv_positions": kvpos / max(seqsteps, 1),
        "algorithmic_bytes_per_launch": by / reps, "achieved_gbs": achieved, "frac": achieved / peak,
    }


def cfg3_strong(dev, rank, world, barrier):
    """BASELINE config 3: a fixed set of 120 ragged utterances with natural EOS, sharded over the ranks (strong scaling)."""
    from gpt_sovits_b200 import shard
    c = CFG3
    sd = synthetic.make_state_dict(seed=c["weight_seed"], eos_scale=c["eos_scale"])
    eng = gsb.T2SEngine(synthetic.S1V2_CONFIG, device=dev)
    eng.load_state_dict(sd, pe=synthetic.sine_pe())
    L = synthetic.config_lens(c["total"], c["lo"], c["hi"], seed=300)
    ids, lens, prompt, bert = synthetic.make_inputs(c["total"], L, c["prompt"], seed=301)
    ids = [t.to(dev) for t in ids]
    bert = [t.to(dev) for t in bert]
    prompt = prompt.to(dev)
    mine_count = [0]

    def infer_chunked(indices):
        """one t2s_generate call: chunks of <= 56 utterances, each decoded in lock-step until its LAST sequence retires"""
        mine_count[0] = len(indices)
        r = eng.infer([ids[i] for i in indices], [bert[i] for i in indices], prompt[indices], top_k=c["top_k"], top_p=c["top_p"],
                      temperature=c["temperature"], repetition_penalty=c["repetition_penalty"], early_stop_num=c["cap"],
                      eos_suppress_steps=c["eos_window"], max_steps=1500, seed=77, utt_ids=list(indices))
        return r.sequences(), r.idx

    def infer_continuous(indices):
        """continuous batching: ONE resident session of <= 56 slots; a retired utterance's slot (and K/V pages) goes to the next
        waiting utterance, so the long utterances of the whole share decode side by side instead of chunk after chunk"""
        mine_count[0] = len(indices)
        sess = gsb.StreamingSession(eng, slots=min(c["slots"], len(indices)), positions=c["hi"] + c["prompt"] + c["cap"] + 8,
                                    slice_steps=c["slice_steps"], admit_min=c["admit_min"], top_k=c["top_k"], top_p=c["top_p"],
                                    temperature=c["temperature"], repetition_penalty=c["repetition_penalty"], early_stop_num=c["cap"],
                                    eos_suppress_steps=c["eos_window"], max_steps=1500, seed=77)
        for i in indices:  # keys = submission order = position in `indices`
            sess.submit([ids[i]], [bert[i]], prompt[i:i + 1])
        ys, ks = [None] * len(indices), [0] * len(indices)
        for key, toks, k in sess:
            ys[key], ks[key] = toks, int(k)
        return ys, ks

    def digest_of(y_list):
        import hashlib
        return hashlib.sha1(b"".join(np.asarray(y, dtype=np.int64).tobytes() for y in y_list)).hexdigest()[:16]

    def timed(fn):
        shard.sharded_infer(fn, L, rank=rank, world=world)  # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        y_list, idx_list = shard.sharded_infer(fn, L, rank=rank, world=world)
        e1.record()
        barrier()
        return max(e0.elapsed_time(e1), 0.0), 1000.0 * (time.perf_counter() - t0), y_list, idx_list

    ms_ch, wall_ch, y_ch, idx_ch = timed(infer_chunked)
    st0 = eng.stats()
    ms, wall, y_list, idx_list = timed(infer_continuous)
    st = eng.stats()
    my_ms = (st["prefill_ms"] + st["decode_ms"]) - (st0["prefill_ms"] + st0["decode_ms"]) if st["decode_ms"] >= st0["decode_ms"] else st["prefill_ms"] + st["decode_ms"]
    eng.close()
    parts = shard.partition_utterances(L, world)
    toks_by_rank = [sum(idx_list[i] for i in p) for p in parts]
    digest = digest_of(y_list)
    same = sum(1 for a, b in zip(y_list, y_ch) if len(a) == len(b) and bool((np.asarray(a) == np.asarray(b)).all()))
    return {"ms": ms, "wall_ms": wall, "chunked_ms": ms_ch, "chunked_digest": digest_of(y_ch), "chunked_tokens": int(sum(idx_ch)), "same_utts": same, "tokens": int(sum(idx_list)), "digest": digest, "idx_mean": float(np.mean(idx_list)), "idx_max": int(max(idx_list)),
            "idx_min": int(min(idx_list)), "utterances_per_rank": [len(p) for p in parts], "tokens_per_rank": toks_by_rank,
            "my_engine_ms": my_ms, "my_utterances": mine_count[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_b32", choices=sorted(WORKLOADS))
    ap.add_argument("--decode-mode", type=int, default=5, help="5 = auto (cluster-stream kernel when the batch fits, else grid-wide phases)")
    ap.add_argument("--cpu-budget", type=float, default=110.0, help="seconds the whole --impl reference run may take (samples are sized to fit)")
    ap.add_argument("--cpu-sample-s", type=float, default=15.0, help="seconds of CPU work of the cpu_baseline leg of our arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the sweep of the other configs (N = 1 only)")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the config-3 strong-scaling job")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.warmup < 3:
        print(f"warning: warmup {args.warmup} < 3 (timing rules ask for >= 3)", file=sys.stderr)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    w = WORKLOADS[name]
    sd = synthetic.make_state_dict(seed=0, eos_scale=0.0)
    eng = gsb.T2SEngine(synthetic.S1V2_CONFIG, device=dev)
    eng.load_state_dict(sd, pe=synthetic.sine_pe())
    from gpt_sovits_b200 import _lib
    eng.set_option(_lib.OPT_DECODE_MODE, args.decode_mode)

    L, ids_h, lens, prompt_h, bert_h = workload_inputs(w, rank)
    ids_d = [t.to(dev) for t in ids_h]
    bert_d = [t.to(dev) for t in bert_h]
    prompt_d = prompt_h.to(dev)
    ids_p = [t.pin_memory() for t in ids_h]
    bert_p = [t.contiguous().pin_memory() for t in bert_h]
    prompt_p = prompt_h.contiguous().pin_memory()
    kw = dict(top_k=w["top_k"], top_p=w["top_p"], temperature=w["temperature"],
              repetition_penalty=w["repetition_penalty"], early_stop_num=w["cap"], eos_suppress_steps=w["eos_window"],
              max_steps=1500)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run(host_io, n, seed0):
        toks, dec_ms, pre_ms, bytes_alg, launches0 = 0, 0.0, 0.0, 0.0, eng.stats()["kernel_launches"]
        for i in range(n):
            if host_io:
                r = eng.infer(ids_p, bert_p, prompt_p, seed=seed0 + i, host_io=True, **kw)
            else:
                r = eng.infer(ids_d, bert_d, prompt_d, seed=seed0 + i, **kw)
            toks += sum(max(v, 0) for v in r.idx)
            st = r.stats
            dec_ms += st["decode_ms"]
            pre_ms += st["prefill_ms"]
            bytes_alg += algorithmic_bytes(st)
        return toks, dec_ms, pre_ms, bytes_alg, eng.stats()["kernel_launches"] - launches0

    # warm-up (both forms)
    run(False, args.warmup, 1000)
    run(True, min(args.warmup, 2), 2000)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- device-resident timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    toks, dec_ms, pre_ms, bytes_alg, launches = run(False, args.steps, 3000)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- end-to-end timing (host buffers)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    f0.record()
    toks_e, _, _, _, _ = run(True, args.steps, 3000)
    f1.record()
    barrier()
    ms_e = max(f0.elapsed_time(f1), 1000.0 * (time.perf_counter() - t_wall))
    clk = clocks.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, ms_e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e = float(t[0]), float(t[1])
        s = torch.tensor([toks, toks_e, launches], device=dev, dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        toks, toks_e, launches = int(s[0]), int(s[1]), int(s[2])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    st_mode = eng.stats()["decode_mode"]

    # ---- the other BASELINE configs on this GPU (N = 1): driver-visible numbers for configs 1, 4, 5
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        for nm in SWEEP:
            try:
                sweep.append(sweep_point(eng, nm, dev, peak))
            except Exception as ex:  # a sweep point must never take the headline down
                sweep.append({"workload": nm, "error": str(ex)[:300]})
    eng.close()

    # ---- BASELINE config 3: fixed total work over N ranks (strong scaling)
    cfg3 = None
    if not args.no_cfg3:
        try:
            c3 = cfg3_strong(dev, rank, world, barrier)
            t3 = torch.tensor([c3["ms"], c3["my_engine_ms"], c3["chunked_ms"]], device=dev, dtype=torch.float64)
            tmax, tmin = t3.clone(), t3.clone()
            if world > 1:
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            job_ms, chunked_ms = float(tmax[0]), float(tmax[2])
            cfg3 = {
                "workload": f"cfg3: {CFG3['total']} ragged utterances ({CFG3['lo']}..{CFG3['hi']} phonemes + {CFG3['prompt']} prompt tokens), "
                            f"natural EOS (EOS row of the head x{CFG3['eos_scale']}, weight seed {CFG3['weight_seed']}), greedy top_k={CFG3['top_k']} "
                            f"rp={CFG3['repetition_penalty']}, on-device retirement; FIXED total work sharded by shard.sharded_infer",
                "scaling": "strong", "n_gpus": world, "tokens": c3["tokens"], "tokens_sha1": c3["digest"], "job_ms": job_ms,
                "tokens_per_s": c3["tokens"] / (job_ms / 1000.0),
                "schedule": f"continuous batching (StreamingSession: one resident session of <= {CFG3['slots']} slots per GPU, slices of "
                            f"{CFG3['slice_steps']} steps, retired slots re-admitted {CFG3['admit_min']} at a time)",
                "chunked": {"job_ms": chunked_ms, "tokens_per_s": c3["chunked_tokens"] / (chunked_ms / 1000.0), "tokens_sha1": c3["chunked_digest"],
                            "schedule": "one t2s_generate call per GPU: chunks of <= 56 utterances, each decoded until its last sequence retires",
                            "utterances_with_the_same_tokens_as_continuous": f"{c3['same_utts']}/{CFG3['total']}",
                            "note": "greedy decoding of RANDOM-INIT weights (near-flat logits): which warp merges which K/V pages of a "
                                    "sequence depends on the other sequences in its cluster, i.e. the fp32 summation order of the "
                                    "attention merge depends on the batch composition, and a 1-ulp difference flips near-ties; the "
                                    "same holds between either schedule and a batch-1 run (scripts/check_cb.py)"},
                "tokens_per_utterance": {"mean": c3["idx_mean"], "min": c3["idx_min"], "max": c3["idx_max"]},
                "utterances_per_rank": c3["utterances_per_rank"], "tokens_per_rank": c3["tokens_per_rank"],
                "engine_ms_slowest_rank": float(tmax[1]), "engine_ms_fastest_rank": float(tmin[1]),
                "limiter": "the job cannot end before its LONGEST utterance (its steps x the step time of a nearly empty batch: the "
                           "latency-bound chain costs about as much at batch 5 as at batch 32); continuous batching removes the "
                           "chunk-after-chunk serialisation at N <= 2 (> 56 utterances per GPU), at N >= 4 a GPU's share fits one "
                           "session and both schedules coincide",
            }
        except Exception as ex:
            cfg3 = {"error": str(ex)[:300]}

    if rank == 0:
        value = toks / (ms / 1000.0)
        e2e = toks_e / (ms_e / 1000.0)
        achieved = bytes_alg / (dec_ms / 1000.0) / 1e9  # rank 0's persistent decode kernel
        traffic = None  # dram bytes of the same launch from the committed ncu --set full capture (profiles/)
        for tf in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", tf)))
                if tr.get("workload") == name and int(tr.get("decode_mode", 1)) == int(st_mode):
                    traffic = float(tr["dram_bytes_read"]) + float(tr["dram_bytes_write"])
                    break
            except Exception:
                pass
        h2d = sum(t.numel() * t.element_size() for t in ids_p) + sum(t.numel() * t.element_size() for t in bert_p) \
            + prompt_p.numel() * prompt_p.element_size()
        d2h = w["batch"] * (w["prompt"] + 1500) * 8 + w["batch"] * 4
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 weights/KV/GEMM operands, f32 accumulate+residual+LN+softmax", "data": "synthetic",
            "config": config_dict(w, name, world),
            "rtf": TOKENS_PER_SEC_AUDIO / (value / (w["batch"] * world)),
            "rtf_note": "per-utterance: seconds of compute per second of audio (25 tokens = 1 s) with the batch decoding in lock-step",
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {
                "kernel": {4: "k_decode_cluster", 1: "k_decode_persistent", 6: "k_decode_wide"}.get(int(st_mode), "decode step graph"),
                "decode_mode": int(st_mode),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_alg / args.steps,
                "launch_ms": dec_ms / args.steps,
                "decode_tokens_per_s_rank0": (toks / world) / (dec_ms / 1000.0),
            },
            "prefill_ms_per_step": pre_ms / args.steps,
            "decode_ms_per_step": dec_ms / args.steps,
            "sweep": sweep,
            "cfg3": cfg3,
        }
        if not args.no_cpu and world == 1:
            arm = CpuArm(w, name)
            arm.calibrate(2, args.cpu_sample_s)
            t, s = arm.sample()
            t2, s2 = arm.sample()
            line["cpu_baseline"] = {"value": (t + t2) / (s + s2), "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                    "sample": arm.describe() + f"; 2 samples, {t + t2} tokens in {s + s2:.1f} s"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
