"""Checkpoint converter: reference checkpoints -> a cache of folded, packed, K-major weights.

    python -m artspeech_b200.convert epoch_2nd_00100.pth g_00935000 cache.pt [--no-state]

The reference loads ``state['net']['ArtsSpeech']`` non-strictly (models.py:685-701) and
``torch.load(g_path)['generator']`` followed by ``remove_weight_norm()`` (test.py:70-73) in every process, and so
did this package: ``nn_util.PlanMixin`` folds weight_norm / eval-mode spectral_norm / BatchNorm and re-lays every
contraction out as 16-bit ``[taps][CoutP][CinP]`` on the first forward.  The cache holds the result of that step
(0.36 GB) with a content hash of the source parameters, so a serving process starts from
``Synthesizer.from_cache(path)`` without folding anything.  ``load_cache`` refuses a cache whose format, compute
dtypes or (when the source parameters are supplied) content hash differ.

Layout of the file (``torch.save``): ``{"format", "hash", "dtypes", "model_params", "vocoder_config",
"distribution", "plans": {"acoustic": {module path: plan}, "vocoder": plan}, "state": {...} or None}`` where a plan
is the module's ``_build_plan`` tree with ``ops.PackedConv`` / ``ops.SmallConv`` flattened to plain dicts.
"""
from __future__ import annotations

import hashlib
import sys
from typing import Optional

import torch

from . import checkpoint, nn_util, ops

FORMAT = 1


# ---------------------------------------------------------------------------------------------------
# plan <-> plain python tree
# ---------------------------------------------------------------------------------------------------
def _flatten(o):
    if isinstance(o, ops.PackedConv):
        return {"__packed__": dict(w=o.w.cpu(), bias=None if o.bias is None else o.bias.cpu(), ntaps=o.ntaps, Cin=o.Cin,
                                   Cout=o.Cout, CinP=o.CinP, CoutP=o.CoutP, taps=[tuple(int(v) for v in t) for t in o.taps])}
    if isinstance(o, ops.SmallConv):
        return {"__small__": dict(w=o.w.cpu(), bias=None if o.bias is None else o.bias.cpu(),
                                  taps=[tuple(int(v) for v in t) for t in o.taps])}
    if torch.is_tensor(o):
        return o.detach().cpu()
    if isinstance(o, dict):
        return {k: _flatten(v) for k, v in o.items()}
    if isinstance(o, tuple):
        return {"__tuple__": [_flatten(v) for v in o]}
    if isinstance(o, list):
        return [_flatten(v) for v in o]
    if o is None or isinstance(o, (int, float, str, bool)):
        return o
    raise TypeError(f"convert: cannot serialise a plan entry of type {type(o).__name__}")


def _inflate(o, device):
    if isinstance(o, dict):
        if "__packed__" in o:
            d = o["__packed__"]
            return ops.PackedConv(d["w"].to(device), None if d["bias"] is None else d["bias"].to(device), d["ntaps"],
                                  d["Cin"], d["Cout"], d["CinP"], d["CoutP"], list(d["taps"]))
        if "__small__" in o:
            d = o["__small__"]
            return ops.SmallConv(d["w"].to(device), None if d["bias"] is None else d["bias"].to(device), list(d["taps"]))
        if "__tuple__" in o:
            return tuple(_inflate(v, device) for v in o["__tuple__"])
        return {k: _inflate(v, device) for k, v in o.items()}
    if isinstance(o, list):
        return [_inflate(v, device) for v in o]
    if torch.is_tensor(o):
        return o.to(device)
    return o


def content_hash(*state_dicts) -> str:
    """sha256 over the (name, dtype, shape, bytes) of every tensor of the source state-dicts, in key order."""
    h = hashlib.sha256()
    for sd in state_dicts:
        for k in sorted(sd):
            t = sd[k].detach().cpu().contiguous()
            h.update(k.encode()); h.update(str(t.dtype).encode()); h.update(str(tuple(t.shape)).encode())
            h.update(t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b"")
    return h.hexdigest()


# ---------------------------------------------------------------------------------------------------
# build / save / load
# ---------------------------------------------------------------------------------------------------
@torch.no_grad()
def build_cache(model, generator, include_state: bool = True) -> dict:
    """Fold and pack every contraction of ``model`` (``models.ArtsSpeech``, second stage) and ``generator``
    (``vocoder.Generator``) on the CPU.  The modules are left untouched."""
    cpu = torch.device("cpu")
    for m in model.modules():                       # the dtype the style encoder hands its extractors at run time
        if hasattr(m, "compute_dtype"):
            m.compute_dtype = model.compute_dtype
    acoustic = {}
    for name, m in model.named_modules():
        if isinstance(m, nn_util.PlanMixin):
            acoustic[name] = _flatten(m._build_plan(cpu))
    voc = _flatten(generator._build_plan(cpu))
    sd, gsd = model.state_dict(), generator.state_dict()
    return {"format": FORMAT, "hash": content_hash(sd, gsd),
            "dtypes": {"acoustic": str(model.compute_dtype), "vocoder": str(generator.compute_dtype)},
            "model_params": dict(checkpoint.MODEL_PARAMS), "vocoder_config": dict(generator.h),
            "distribution": {k: v.detach().cpu() for k, v in model.distribution.items()},
            "plans": {"acoustic": acoustic, "vocoder": voc},
            "state": {"acoustic": {k: v.detach().cpu() for k, v in sd.items()},
                      "vocoder": {k: v.detach().cpu() for k, v in gsd.items()}} if include_state else None}


def install_cache(cache: dict, model, generator, device) -> None:
    """Give ``model`` / ``generator`` (already on ``device``) the cached plans instead of folding their own."""
    if cache.get("format") != FORMAT:
        raise ValueError(f"weight cache format {cache.get('format')} != {FORMAT}")
    want = {"acoustic": str(model.compute_dtype), "vocoder": str(generator.compute_dtype)}
    if cache["dtypes"] != want:
        raise ValueError(f"weight cache was packed for {cache['dtypes']}, the modules compute in {want}")
    device = next(model.parameters()).device          # with its index ("cuda" -> "cuda:0"): what plan() compares with
    mods = {name: m for name, m in model.named_modules() if isinstance(m, nn_util.PlanMixin)}
    if set(mods) != set(cache["plans"]["acoustic"]):
        raise ValueError("weight cache does not match the module tree")
    for name, m in mods.items():
        plan = _inflate(cache["plans"]["acoustic"][name], device)
        plan["_device"] = str(device)
        m._plan = plan
    generator._plan = _inflate(cache["plans"]["vocoder"], device)
    nn_util.bump_plan_epoch()


def load_cache(path: str, device="cuda:0", expect_hash: Optional[str] = None):
    """-> (model, generator) on ``device`` with the cached plans installed.  With ``state`` in the cache the modules
    carry the source parameters (``state_dict()`` round-trips); without it they keep their constructor
    initialisation, which inference never reads."""
    from . import models, vocoder
    cache = torch.load(path, map_location="cpu", weights_only=False)
    if expect_hash is not None and cache.get("hash") != expect_hash:
        raise ValueError("weight cache content hash differs from the expected checkpoint hash")
    model = models.ArtsSpeech(checkpoint.AttrDict(cache["model_params"]), stage="second",
                              distribution=cache["distribution"])
    gen = vocoder.Generator(checkpoint.AttrDict(cache["vocoder_config"]))
    gen.remove_weight_norm()
    if cache.get("state") is not None:
        model.load_state_dict(cache["state"]["acoustic"], strict=True)
        gen.load_state_dict(cache["state"]["vocoder"], strict=True)
    model, gen = model.to(device).eval(), gen.to(device).eval()
    model.distribution = {k: v.to(device) for k, v in model.distribution.items()}
    install_cache(cache, model, gen, device)
    return model, gen


def convert(acoustic_ckpt: str, generator_ckpt: str, out_path: str, include_state: bool = True) -> str:
    """The CLI: reference checkpoint files -> ``out_path``.  Returns the content hash."""
    from . import models, vocoder
    state = torch.load(acoustic_ckpt, map_location="cpu", weights_only=False)
    params = state["net"]["ArtsSpeech"] if "net" in state else state          # models.py:686-690
    model = models.ArtsSpeech(checkpoint.AttrDict(checkpoint.MODEL_PARAMS), stage="second",
                              distribution=checkpoint.default_distribution())
    model.load_state_dict(params, strict=False)                                 # non-strict, as the reference
    gstate = torch.load(generator_ckpt, map_location="cpu", weights_only=False)
    gen = vocoder.Generator(checkpoint.AttrDict(checkpoint.VOCODER_CONFIG))
    gen.load_state_dict(gstate["generator"] if "generator" in gstate else gstate)   # test.py:70-71
    gen.remove_weight_norm()                                                    # test.py:73
    cache = build_cache(model.eval(), gen.eval(), include_state)
    torch.save(cache, out_path)
    return cache["hash"]


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if len(args) != 3:
        raise SystemExit(__doc__)
    print(convert(args[0], args[1], args[2], include_state="--no-state" not in sys.argv))
