"""Pin oracle/restate.py (the travelling CPU restatement) to the reference: against the committed
fixtures everywhere, and against the unmodified reference modules where /root/reference exists."""
import os

import pytest
import torch

from oracle import ref_loader, restate
from tests import util


@pytest.mark.parametrize("case", ["a_pred_dur", "b_forced_dur", "c_short"])
def test_restatement_matches_reference_goldens(case):
    g = util.load_golden("acoustic_small.pt")
    model = util.acoustic_model(g["checkpoint_seed"])
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    c = g["cases"][case]
    dist = {k: v.cpu() for k, v in model.distribution.items()}
    out, aux = restate.artsspeech_test(sd, c["tokens"], c["ref_mel"], dist, durations=None if c["durations"] is None
                                       else c["durations"].view(-1), want_aux=True)
    o = c["out"]
    assert torch.equal(aux["pred_dur"].view(-1), o["pred_dur"].view(-1))
    assert (aux["style"] - o["style"]).abs().max().item() < 1e-5
    for k in ("F0", "N", "EMA", "f0_ext", "n_ext", "ema_ext"):
        assert (aux[k] - o[k]).abs().max().item() < 1e-5, k
    assert (out - o["mel"]).abs().max().item() < 5e-5


@pytest.mark.reference
@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")
def test_restatement_matches_live_reference():
    from artspeech_b200 import checkpoint
    from oracle import ref_runner
    dist = ref_loader.load_distribution()
    ours = checkpoint.build_random_artsspeech(3, dist)
    ref = ref_loader.build_reference_artsspeech(dist)
    ref.load_state_dict(ours.state_dict(), strict=True)      # identical parameter tree
    g = torch.Generator().manual_seed(11)
    tok = torch.randint(1, 178, (1, 23), generator=g)
    mel = torch.randn(1, 80, 97, generator=g) * 0.5
    r = ref_runner.reference_test_step(ref, tok, mel)
    out, aux = restate.artsspeech_test(ours.state_dict(), tok, mel, dist, want_aux=True)
    assert torch.equal(aux["pred_dur"].view(-1), r["pred_dur"].view(-1))
    assert (out - r["mel"]).abs().max().item() < 5e-5
    gen = checkpoint.build_random_generator(3)
    refg = ref_loader.build_reference_generator()
    refg.remove_weight_norm()
    refg.load_state_dict(gen.state_dict(), strict=True)
    with torch.no_grad():
        w = refg(r["mel"])
    assert (restate.generator_forward(gen.state_dict(), r["mel"]) - w).abs().max().item() < 1e-5
