"""ctypes binding of libt2s_b200.so (include/t2s_b200.h).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libt2s_b200.so")
import glob  # noqa: E402

# engine.cu is the translation unit; every header it includes (and the C-ABI header) is a rebuild dependency
SOURCES = [os.path.join(_HERE, "csrc", "engine.cu")] + sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cuh"))) + \
          [os.path.join(os.path.dirname(_HERE), "include", "t2s_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

F32, F16, BF16 = 0, 1, 2

# tensor ids (enum order of t2s_b200.h)
(W_BERT_PROJ_W, W_BERT_PROJ_B, W_TEXT_EMB, W_TEXT_ALPHA, W_AUDIO_EMB, W_AUDIO_ALPHA, W_PE, W_PREDICT,
 W_IN_PROJ_W, W_IN_PROJ_B, W_OUT_PROJ_W, W_OUT_PROJ_B, W_LIN1_W, W_LIN1_B, W_LIN2_W, W_LIN2_B,
 W_NORM1_W, W_NORM1_B, W_NORM2_W, W_NORM2_B, W_COUNT) = range(21)

OPT_DECODE_MODE, OPT_PREFILL_GEMM, OPT_NUM_CTAS, OPT_CHECK_STEPS, OPT_TC_DECODE_MIN_BATCH = 0, 1, 2, 3, 4
OPT_SESSION_SLOTS, OPT_SESSION_POSITIONS, OPT_HOOKS_BY_UTTERANCE = 5, 6, 7

EXPORTS = ["t2s_create", "t2s_destroy", "t2s_last_error", "t2s_load_tensor", "t2s_prefill", "t2s_admit", "t2s_release_slots", "t2s_set_utterance_ids", "t2s_decode",
           "t2s_result", "t2s_generate", "t2s_set_forced_tokens", "t2s_set_logits_capture",
           "t2s_get_sampled", "t2s_set_option", "t2s_get_stats", "t2s_sampler_test", "t2s_bench_barrier", "t2s_set_timeline",
           "t2s_codes_to_latent"]


class ModelConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layer", "d_model", "n_head", "d_ff", "vocab", "phoneme_vocab",
                                         "bert_dim", "eos", "pe_len", "max_batch")]


class Request(C.Structure):
    _fields_ = [
        ("batch", C.c_int32),
        ("phoneme_ids", C.c_void_p),
        ("phoneme_lens", C.POINTER(C.c_int32)),
        ("bert", C.POINTER(C.c_void_p)),
        ("bert_stride_c", C.POINTER(C.c_int64)),
        ("bert_stride_t", C.POINTER(C.c_int64)),
        ("bert_dtype", C.c_int32),
        ("prompt", C.c_void_p),
        ("prompt_row_stride", C.c_int64),
        ("prompt_len", C.c_int32),
        ("top_k", C.c_int32),
        ("top_p", C.c_float),
        ("temperature", C.c_float),
        ("repetition_penalty", C.c_float),
        ("early_stop_num", C.c_int32),
        ("eos_suppress_steps", C.c_int32),
        ("max_steps", C.c_int32),
        ("seed", C.c_uint64),
        ("inputs_on_host", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("prefill_ms", C.c_double), ("decode_ms", C.c_double),
        ("decode_steps", C.c_int64), ("decode_tokens", C.c_int64), ("decode_kv_positions", C.c_int64),
        ("kernel_launches", C.c_int64), ("prefill_rows", C.c_int64),
        ("weight_bytes_per_step", C.c_int64), ("kv_bytes_per_position", C.c_int64),
        ("num_sms", C.c_int32), ("decode_mode", C.c_int32),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(s) for s in SOURCES if os.path.exists(s))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, SOURCES[0]]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Loads libt2s_b200.so.  There is deliberately no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  gpt-sovits_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.t2s_create.argtypes = [C.POINTER(ModelConfig), C.POINTER(vp)]
    lib.t2s_destroy.argtypes = [vp]
    lib.t2s_destroy.restype = None
    lib.t2s_last_error.argtypes = []
    lib.t2s_last_error.restype = C.c_char_p
    lib.t2s_load_tensor.argtypes = [vp, i32, i32, vp, i32, i64, i32, vp]
    lib.t2s_prefill.argtypes = [vp, C.POINTER(Request), vp]
    lib.t2s_admit.argtypes = [vp, C.POINTER(Request), vp]
    lib.t2s_release_slots.argtypes = [vp, C.POINTER(i32), i32, vp]
    lib.t2s_set_utterance_ids.argtypes = [vp, C.POINTER(i32), i32]
    lib.t2s_decode.argtypes = [vp, i32, vp, C.POINTER(i32)]
    lib.t2s_result.argtypes = [vp, vp, i64, i32, C.POINTER(i32), vp]
    lib.t2s_generate.argtypes = [vp, C.POINTER(Request), vp, i64, i32, C.POINTER(i32), vp]
    lib.t2s_set_forced_tokens.argtypes = [vp, vp, i32]
    lib.t2s_set_logits_capture.argtypes = [vp, vp, i32]
    lib.t2s_get_sampled.argtypes = [vp, vp, i32, vp]
    lib.t2s_set_option.argtypes = [vp, i32, i64]
    lib.t2s_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.t2s_sampler_test.argtypes = [vp, vp, i32, i32, vp, i32, i32, C.c_float, C.c_float, C.c_float, C.c_uint64, i32,
                                     vp, vp, vp]
    lib.t2s_bench_barrier.argtypes = [vp, i32, i32, C.POINTER(C.c_float), vp]
    lib.t2s_set_timeline.argtypes = [vp, vp, i32, i32]
    lib.t2s_codes_to_latent.argtypes = [vp, vp, i32, vp, i32, i32, i32, vp, vp]
    for name in EXPORTS:
        if name not in ("t2s_destroy", "t2s_last_error"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("t2s_b200: " + load().t2s_last_error().decode("utf-8", "replace"))
