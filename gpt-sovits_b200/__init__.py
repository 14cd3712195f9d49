"""B200-native text-to-semantic (T2S) autoregressive decode for GPT-SoVITS.

Drop-in for ``Text2SemanticDecoder.infer_panel`` / ``infer_panel_batch_infer`` /
``infer_panel_naive`` / ``infer_panel_naive_batched`` (reference:
GPT_SoVITS/AR/models/t2s_model.py:583-935).  All compute runs in hand-written sm_100a CUDA kernels
behind the C-ABI library ``libt2s_b200.so`` (include/t2s_b200.h); there is no CPU fallback: creating an
engine without the library or without a CUDA device raises.
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic"]
