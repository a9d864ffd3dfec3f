"""Checkpoint converter on the CPU: reference-format checkpoint files -> weight cache; the cached
plans are exactly what the modules fold for themselves, the content hash pins the source parameters."""
import torch

from artspeech_b200 import checkpoint, convert, ops
from tests import util


def _same(a, b, path=""):
    if isinstance(a, (ops.PackedConv, ops.SmallConv)):
        assert type(a) is type(b), path
        assert torch.equal(a.w.cpu(), b.w.cpu()), path
        assert (a.bias is None) == (b.bias is None) and (a.bias is None or torch.equal(a.bias.cpu(), b.bias.cpu())), path
        assert [tuple(t) for t in a.taps] == [tuple(t) for t in b.taps], path
        if isinstance(a, ops.PackedConv):
            assert (a.ntaps, a.Cin, a.Cout, a.CinP, a.CoutP) == (b.ntaps, b.Cin, b.Cout, b.CinP, b.CoutP), path
    elif torch.is_tensor(a):
        assert torch.equal(a.cpu(), b.cpu()), path
    elif isinstance(a, dict):
        assert set(a) == set(b), path
        for k in a:
            _same(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert type(a) is type(b) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a == b, path


def test_convert_round_trip(tmp_path):
    model = util.acoustic_model(0)
    gen_wn = checkpoint.build_random_generator(0, remove_wn=False)          # as shipped: weight-normed (test.py:70-73)
    # files in the reference's formats (models.py:686-690, test.py:70-71)
    a_path, g_path, out = tmp_path / "epoch_2nd.pth", tmp_path / "g_0", tmp_path / "cache.pt"
    torch.save({"net": {"ArtsSpeech": model.state_dict()}, "epoch": 1, "iters": 2}, a_path)
    torch.save({"generator": gen_wn.state_dict()}, g_path)
    digest = convert.convert(str(a_path), str(g_path), str(out))
    cache = torch.load(out, map_location="cpu", weights_only=False)
    assert cache["format"] == convert.FORMAT and cache["hash"] == digest
    # the same parameters hash the same, different ones do not
    gen = checkpoint.build_random_generator(0)
    assert digest == convert.content_hash(model.state_dict(), gen.state_dict())
    assert digest != convert.content_hash(model.state_dict(), checkpoint.build_random_generator(1).state_dict())
    # cached plans == what the modules fold for themselves (bit for bit)
    cpu = torch.device("cpu")
    n = 0
    for name, m in model.named_modules():
        if hasattr(m, "_build_plan") and name in cache["plans"]["acoustic"]:
            _same(convert._inflate(cache["plans"]["acoustic"][name], cpu), m._build_plan(cpu), name)
            n += 1
    assert n == len(cache["plans"]["acoustic"]) >= 8
    _same(convert._inflate(cache["plans"]["vocoder"], cpu), gen._build_plan(cpu), "vocoder")
    # a cache without the source parameters is the packed weights only
    small = convert.build_cache(model, gen, include_state=False)
    assert small["state"] is None and small["hash"] == digest
    util._MODELS.clear()
