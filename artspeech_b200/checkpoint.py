"""Deterministic, well-conditioned random checkpoints (the real ones are not redistributable).

The reference's checkpoints are absent (``/root/reference/.MISSING_LARGE_BLOBS``; LFS pointers),
and a default-initialised reference explodes numerically (old-style ``spectral_norm``
starts from random ``u, v`` and never iterates in eval mode).  These recipes build state-dicts with
the reference's exact key layout whose activations stay O(1), from a seed, on any machine:
  * ``randomize_generator``  — HiFi-GAN generator with fan-in scaled weights;
  * ``condition_spectral_norm`` — runs the power iteration ``u <- W v, v <- W^T u`` on every
    old-style spectral-norm layer so that ``sigma = u.W v`` is the true top singular value.
Both the oracle and the CUDA path load the *same* state-dict, so parity does not depend on the
recipe; it only has to be reproducible.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


@torch.no_grad()
def randomize_generator(gen: nn.Module, seed: int = 0, gain: float = 1.0) -> nn.Module:
    """Re-initialise a ``vocoder.Generator`` (weight-normed or not) in place, deterministically."""
    g = torch.Generator().manual_seed(seed)
    for name, m in gen.named_modules():
        if not isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
            continue
        v = m.weight_v if hasattr(m, "weight_v") else m.weight
        if isinstance(m, nn.ConvTranspose1d):
            # each output sample sees Cin * k / stride taps
            fan_in = v.shape[0] * v.shape[2] / m.stride[0]
        else:
            fan_in = v.shape[1] * v.shape[2]
        std = gain / math.sqrt(fan_in)
        # residual branches (convs2) are damped so the 3-deep residual chains stay O(1)
        if ".convs2." in name:
            std *= 0.5
        v.copy_(torch.randn(v.shape, generator=g) * std)
        if hasattr(m, "weight_g"):
            n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(m.weight_g.shape)
            m.weight_g.copy_(n)
        if m.bias is not None:
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    if hasattr(gen, "invalidate_plan"):
        gen.invalidate_plan()
    return gen


@torch.no_grad()
def condition_spectral_norm(model: nn.Module, iters: int = 40) -> int:
    """Power-iterate ``weight_u`` / ``weight_v`` of every old-style spectral_norm layer in place."""
    n = 0
    for m in model.modules():
        if hasattr(m, "weight_orig") and hasattr(m, "weight_u") and hasattr(m, "weight_v"):
            w = m.weight_orig.detach().float()
            wm = w.reshape(w.shape[0], -1)
            u, v = m.weight_u.float().clone(), m.weight_v.float().clone()
            for _ in range(iters):
                v = torch.nn.functional.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
                u = torch.nn.functional.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
            m.weight_u.copy_(u)
            m.weight_v.copy_(v)
            n += 1
    return n


@torch.no_grad()
def randomize_batchnorm(model: nn.Module, seed: int = 0) -> int:
    """Give every BatchNorm non-trivial running statistics / affine so folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    for m in model.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75)
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            n += 1
    return n


@torch.no_grad()
def condition_artsspeech(model: nn.Module, seed: int = 0) -> nn.Module:
    """Turn a freshly constructed ``models.ArtsSpeech`` (either implementation: the parameter trees
    are identical) into a numerically sane random checkpoint:
      * spectral-norm ``u, v`` power-iterated,
      * BatchNorm statistics randomised,
      * the zero-initialised prenet projections (RelTransformerEnc.py:313-314) made non-zero so the
        prenet actually contributes.
    Deterministic given the construction seed and ``seed``."""
    g = torch.Generator().manual_seed(seed + 1000)
    condition_spectral_norm(model, iters=40)
    randomize_batchnorm(model, seed)
    for name, p in model.named_parameters():
        if name.endswith("pre.proj.weight"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        elif name.endswith("pre.proj.bias"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    for m in model.modules():
        if hasattr(m, "invalidate_plan"):
            m.invalidate_plan()
    return model


MODEL_PARAMS = dict(hidden_dim=512, n_token=178, style_dim=256, n_layer=3, dim_in=64, max_conv_dim=512,
                    n_mels=80, dropout=0.2)   # Configs/config.yaml:30-38

VOCODER_CONFIG = dict(resblock="1", upsample_rates=[10, 5, 3, 2], upsample_kernel_sizes=[20, 10, 6, 4],
                      upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                      resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80,
                      sampling_rate=24000, hop_size=300)   # Vocoder/config.json:2,12-25

# Data/stats.json entries [2], [3] (mean, std) exactly as test.py:49-56 reads them; committed here
# because /root/reference does not exist on the GPU box.
def default_distribution(device="cpu"):
    from . import _stats
    return {k: torch.tensor(v, dtype=torch.float32, device=device) for k, v in _stats.DISTRIBUTION.items()}


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def build_random_artsspeech(seed: int = 0, distribution=None):
    """Seeded, conditioned ``artspeech_b200.models.ArtsSpeech(stage='second')`` on CPU (eval)."""
    from . import models
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        m = models.ArtsSpeech(AttrDict(MODEL_PARAMS), stage="second",
                              distribution=distribution if distribution is not None else default_distribution())
    condition_artsspeech(m, seed)
    return m.eval()


def build_random_generator(seed: int = 0, remove_wn: bool = True):
    """Seeded ``artspeech_b200.vocoder.Generator`` with O(1) activations (eval, weight norm removed
    like test.py:73 does)."""
    from . import vocoder
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        g = vocoder.Generator(AttrDict(VOCODER_CONFIG))
    randomize_generator(g, seed)
    if remove_wn:
        g.remove_weight_norm()
    return g.eval()
