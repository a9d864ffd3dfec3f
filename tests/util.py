import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances of BASELINE.json north_star
MEL_MAX_ABS = 1e-2
MEL_MEAN_ABS = 1e-3
WAV_SNR_DB = 35.0


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def snr_db(x, ref):
    x, ref = x.double().flatten(), ref.double().flatten()
    return float(10 * torch.log10((ref ** 2).sum() / ((x - ref) ** 2).sum().clamp_min(1e-300)))


def mas_cases():
    z = np.load(os.path.join(GOLDEN, "mas.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in z.files})
    for key in keys:
        v = z[key + "_value"]
        n = v.size
        unpack = lambda a: np.unpackbits(a)[:n].reshape(v.shape).astype(np.float32)
        yield key, v, z[key + "_xlen"], z[key + "_ylen"], unpack(z[key + "_path1"]), unpack(z[key + "_path2"])


_MODELS = {}


def acoustic_model(seed=0):
    """Seeded conditioned ArtsSpeech (CPU, eval); cached because construction takes seconds."""
    from artspeech_b200 import checkpoint
    if ("a", seed) not in _MODELS:
        _MODELS[("a", seed)] = checkpoint.build_random_artsspeech(seed)
    return _MODELS[("a", seed)]


def generator(seed=0):
    from artspeech_b200 import checkpoint
    if ("g", seed) not in _MODELS:
        _MODELS[("g", seed)] = checkpoint.build_random_generator(seed)
    return _MODELS[("g", seed)]
