"""The C-ABI library loads without a GPU and exports every symbol include/artspeech_b200.h declares;
compute entry points fail loudly (no CPU fallback)."""
import os
import re

import pytest
import torch

from artspeech_b200 import _lib, mas, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "artspeech_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(as_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.as_version() >= 100
    assert lib.as_conv_tile_n(640) == ops.conv_tile_n(640) == 128
    for c in (1, 16, 17, 32, 80, 128, 192, 256, 512, 640, 1024, 1536, 2560):
        assert lib.as_conv_tile_n(c) == ops.conv_tile_n(c)
    # direction bits of 64x200x1000 fit in shared memory (no workspace); long items spill to the caller's workspace
    assert lib.as_mas_workspace_bytes(64, 200, 1000) == 0
    assert lib.as_mas_workspace_bytes(2, 1280, 4000) >= 2 * 1280 * 4000 // 8
    assert lib.as_mas_workspace_bytes(0, 200, 1000) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(_lib.AsError):
        mas.maximum_path_lens(torch.zeros(1, 4, 8), torch.tensor([4]), torch.tensor([8]))
    with pytest.raises(_lib.AsError):
        ops.layernorm(torch.zeros(1, 4, 8), torch.ones(8), torch.zeros(8), 1e-5, out_a=torch.float32)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "artspeech_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
