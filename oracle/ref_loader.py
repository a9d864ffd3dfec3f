"""TEST INFRASTRUCTURE — loader for the *unmodified* reference (Zhongxu-Wang/ArtSpeech).

Only usable where ``/root/reference`` exists (the build container).  It is used to
  * pin ``oracle/restate.py`` (the travelling CPU restatement) against the real reference, and
  * generate the committed fixtures under ``tests/golden`` (see ``oracle/make_golden.py``).
Nothing in the product package imports this file.

Patches applied (all oracle-side; the reference tree is never written):
  1. ``sys.path`` shims for ``munch`` / ``attrdict`` / ``matplotlib`` (models.py:6,
     Vocoder/vocoder_utils.py:3-7 import them; none is installed here).
  2. ``torch.load`` returns seeded random state-dicts for ``Utils/JDC/bst.t7`` and
     ``Utils/EMA/200000.pth.tar`` (models.py:378,382) — both blobs are absent from the checkout
     (``.MISSING_LARGE_BLOBS``).  The values are irrelevant: every oracle run overwrites the whole
     model with a checkpoint produced by ``artspeech_b200.checkpoint``.
  3. Without a GPU, ``.to("cuda")`` is mapped to a no-op (models.py:367-368,377,381 hard-code it).
"""
from __future__ import annotations

import contextlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("ARTSPEECH_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models.py"))


def _is_cuda_arg(a) -> bool:
    if isinstance(a, str):
        return a.startswith("cuda")
    if isinstance(a, torch.device):
        return a.type == "cuda"
    return False


@contextlib.contextmanager
def reference_env():
    """Context in which the reference modules import and construct on a CPU-only box."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    added = [p for p in (_SHIMS, REFERENCE_ROOT) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    orig_load = torch.load
    orig_mod_to = torch.nn.Module.to
    orig_t_to = torch.Tensor.to
    have_cuda = torch.cuda.is_available()

    def fake_load(f, *a, **kw):
        name = str(f)
        if name.endswith("Utils/JDC/bst.t7"):
            from Utils.JDC.model import JDCNet
            g = torch.Generator().manual_seed(11)
            with torch.random.fork_rng():
                torch.manual_seed(11)
                sd = JDCNet(num_class=1, seq_len=192).state_dict()
            del g
            return {"net": sd}
        if name.endswith("Utils/EMA/200000.pth.tar"):
            from Utils.EMA.EMA_Predictor import EMA_Predictor
            with torch.random.fork_rng():
                torch.manual_seed(12)
                sd = EMA_Predictor().state_dict()
            return {"model": sd}
        return orig_load(f, *a, **kw)

    def mod_to(self, *a, **kw):
        if not have_cuda and ((a and _is_cuda_arg(a[0])) or _is_cuda_arg(kw.get("device"))):
            return self
        return orig_mod_to(self, *a, **kw)

    def t_to(self, *a, **kw):
        if not have_cuda and ((a and _is_cuda_arg(a[0])) or _is_cuda_arg(kw.get("device"))):
            return self
        return orig_t_to(self, *a, **kw)

    torch.load = fake_load
    torch.nn.Module.to = mod_to
    torch.Tensor.to = t_to
    try:
        yield
    finally:
        torch.load = orig_load
        torch.nn.Module.to = orig_mod_to
        torch.Tensor.to = orig_t_to
        for p in added:
            if p in sys.path:
                sys.path.remove(p)


def load_distribution(device="cpu"):
    """Normalisation constants exactly as test.py:49-56,75-79 reads them from Data/stats.json."""
    import json
    with open(os.path.join(REFERENCE_ROOT, "Data", "stats.json")) as f:
        data = json.load(f)
    out = {}
    for key in ("EMA", "pitch", "energy"):
        _, _, mean_val, std_val = data[key]
        out[f"{key}_mean"] = torch.tensor(mean_val).to(device)
        out[f"{key}_std"] = torch.tensor(std_val).to(device)
    return out


def build_reference_artsspeech(distribution=None):
    """``models.ArtsSpeech(stage='second')`` of the reference with config.yaml:30-38 params."""
    with reference_env():
        import yaml
        import models as ref_models
        from munch import Munch
        cfg = yaml.safe_load(open(os.path.join(REFERENCE_ROOT, "Configs", "config.yaml")))
        args = Munch(cfg["model_params"])
        if distribution is None:
            distribution = load_distribution()
        with torch.random.fork_rng():
            torch.manual_seed(0)
            m = ref_models.ArtsSpeech(args, stage="second", distribution=distribution)
        return m.eval()


def build_reference_generator():
    """``Vocoder.vocoder.Generator(AttrDict(config.json))`` of the reference (test.py:66-69)."""
    with reference_env():
        import json
        from attrdict import AttrDict
        from Vocoder.vocoder import Generator
        h = AttrDict(json.load(open(os.path.join(REFERENCE_ROOT, "Vocoder", "config.json"))))
        with torch.random.fork_rng():
            torch.manual_seed(0)
            g = Generator(h)
        return g.eval()


def reference_mas():
    with reference_env():
        import S_monotonic_align as sma
        return sma
