"""B200-native text-to-semantic (T2S) autoregressive decode for GPT-SoVITS.

Drop-in for ``Text2SemanticDecoder.infer_panel`` / ``infer_panel_batch_infer`` /
``infer_panel_naive`` / ``infer_panel_naive_batched`` (reference:
GPT_SoVITS/AR/models/t2s_model.py:583-935).  All compute runs in hand-written sm_100a CUDA kernels
behind the C-ABI library ``libt2s_b200.so`` (include/t2s_b200.h); there is no CPU fallback: creating an
engine without the library or without a CUDA device raises.
"""
from . import synthetic  # noqa: F401
from .decoder import (  # noqa: F401
    engine_for,
    infer_panel,
    infer_panel_batch_infer,
    infer_panel_naive,
    infer_panel_naive_batched,
    infer_panel_stream,
    patch_reference,
    unpatch_reference,
)
from .engine import EOS_WINDOW_BATCH, EOS_WINDOW_NAIVE, MAX_STEPS, InferResult, T2SEngine  # noqa: F401
from .checkpoint import engine_from_checkpoint, read_checkpoint  # noqa: F401
from .batching import StreamingSession, bucket_batches, recovery_order  # noqa: F401

__all__ = ["synthetic", "T2SEngine", "InferResult", "patch_reference", "unpatch_reference", "engine_for",
           "infer_panel", "infer_panel_naive", "infer_panel_naive_batched", "infer_panel_batch_infer",
           "MAX_STEPS", "EOS_WINDOW_NAIVE", "EOS_WINDOW_BATCH", "read_checkpoint", "engine_from_checkpoint",
           "StreamingSession", "bucket_batches", "recovery_order", "infer_panel_stream"]
