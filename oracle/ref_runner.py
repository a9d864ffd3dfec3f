"""TEST INFRASTRUCTURE — run the *unmodified* reference modules (container only, needs
/root/reference) on a state-dict produced by ``artspeech_b200.checkpoint`` and return every
intermediate of ``ArtsSpeech.forward(step='test')`` (models.py:356-371) + ``Generator`` output.
Used to pin ``oracle/restate.py`` and to generate ``tests/golden``."""
from __future__ import annotations

import torch

from . import ref_loader as rl


@torch.no_grad()
def reference_test_step(ref_model, tokens, mel, durations=None):
    """Batch-1 reference synthesis.  ``tokens`` int64 [1,Tt], ``mel`` fp32 [1,80,Tr].
    Returns a dict of intermediates (all CPU tensors).  With ``durations`` the reference's own
    lines 362-370 are re-executed with those integer durations (the predictor is bypassed)."""
    cap = {}

    def hook(name):
        def fn(mod, inp, out):
            cap[name] = out
        return fn

    hs = [ref_model.text_encoder.register_forward_hook(hook("T_en")),
          ref_model.arts_encoder.register_forward_hook(hook("A_en")),
          ref_model.style_encoder.register_forward_hook(hook("style_out")),
          ref_model.durationPredictor.register_forward_hook(hook("duration")),
          ref_model.artsPredictor.register_forward_hook(hook("arts")),
          ref_model.style_encoder.pitch_extractor.register_forward_hook(hook("f0_raw")),
          ref_model.style_encoder.ema_extractor.register_forward_hook(hook("ema_raw"))]
    lens = torch.LongTensor([tokens.shape[1]])
    mlens = torch.LongTensor([mel.shape[2]])
    with rl.reference_env():
        mel_out = ref_model([tokens, lens, mel, mlens, None, None, None], None, None, step="test")
    for h in hs:
        h.remove()
    f0_ext, n_ext, ema_ext, style = cap["style_out"]
    out = dict(T_en=cap["T_en"], A_en=cap["A_en"], f0_ext=f0_ext, n_ext=n_ext, ema_ext=ema_ext, style=style,
               duration=cap["duration"], pred_dur=torch.round(cap["duration"].squeeze(0)).clamp(min=1).long(),
               F0=cap["arts"][0], N=cap["arts"][1], EMA=cap["arts"][2], mel=mel_out,
               f0_raw=cap["f0_raw"], ema_raw=cap["ema_raw"])
    if durations is not None:
        d = durations.long().view(-1)
        aln = torch.zeros(tokens.shape[1], int(d.sum()))
        c = 0
        for i in range(aln.size(0)):                      # models.py:363-366
            aln[i, c:c + int(d[i])] = 1
            c += int(d[i])
        T_en = cap["T_en"].transpose(1, 2) @ aln.unsqueeze(0)
        A_en = cap["A_en"].transpose(1, 2) @ aln.unsqueeze(0)
        F0, N, EMA = ref_model.artsPredictor(A_en, style)
        out.update(F0=F0, N=N, EMA=EMA, mel=ref_model.decoder(T_en, style, F0, N, EMA), pred_dur=d)
    return {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in out.items()}
