"""TEST INFRASTRUCTURE — loads the *unmodified* reference decoder from /root/reference on CPU.

Only usable where /root/reference exists (the build container); the GPU box does not have it, so
nothing in `-m gpu` tests, smoke() or bench.py imports this module.  It is used by
oracle/make_goldens.py to produce tests/golden/*.npz and by the container-only tests that pin
oracle/t2s_oracle.py against the live reference.

The one stub: `torchmetrics` (a training-only import at GPT_SoVITS/AR/models/t2s_model.py:9) is not
installed in this image; a dummy module is injected before import.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from typing import Dict, List, Optional

import torch

REFERENCE_ROOT = "/root/reference/GPT_SoVITS"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "AR", "models", "t2s_model.py"))


def _install_stubs() -> None:
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tmc = types.ModuleType("torchmetrics.classification")

        class MulticlassAccuracy:  # training metric, never touched at inference
            def __init__(self, *a, **k):
                pass

        tmc.MulticlassAccuracy = MulticlassAccuracy
        tm.classification = tmc
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.classification"] = tmc


def import_reference():
    """Returns the reference module AR.models.t2s_model."""
    if not reference_available():
        raise RuntimeError("reference not present (expected at %s)" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import AR.models.t2s_model as ref  # noqa

    return ref


def build_reference_model(state_dict: Dict[str, torch.Tensor], config: dict):
    ref = import_reference()
    model = ref.Text2SemanticDecoder(config).eval()
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("ar_accuracy_metric") or k.startswith("loss_fct") for k in missing), missing
    return model


class SampleHook:
    """Replaces the module-global `sample` (looked up at call time, t2s_model.py:714/:891) to record the
    pre-sampling logits of every step and optionally teacher-force the emitted tokens."""

    def __init__(self, forced: Optional[torch.Tensor] = None):
        self.forced = forced  # [B, n_steps] int64 or None
        self.logits: List[torch.Tensor] = []
        self.samples: List[torch.Tensor] = []
        self.step = 0

    def install(self):
        ref = import_reference()
        self._ref = ref
        self._orig = ref.sample

        def hooked(logits, previous_tokens=None, **kw):
            self.logits.append(logits.detach().clone())
            out, probs = self._orig(logits, previous_tokens, **kw)
            self.samples.append(out.detach().clone())
            if self.forced is not None and self.step < self.forced.shape[1]:
                out = self.forced[: out.shape[0], self.step : self.step + 1].to(out.dtype)
            self.step += 1
            return out, probs

        ref.sample = hooked
        return self

    def remove(self):
        self._ref.sample = self._orig


@contextlib.contextmanager
def quiet():
    """The reference prints progress bars and 'T2S Decoding EOS' lines."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield
