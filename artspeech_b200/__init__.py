"""artspeech_b200 — B200-native (sm_100a) implementation of ArtSpeech's batched synthesis forward
pass and MAS, behind the reference's own nn.Module / function API.

Public surface (mirrors the reference):
  * ``artspeech_b200.mas``      — maximum_path1 / maximum_path2 / maximum_path / mask_from_lens
  * ``artspeech_b200.vocoder``  — Generator
  * ``artspeech_b200.models``   — ArtsSpeech, StyleEncoder, DurationPredictor, ArtsPredictor, Decoder, build_model
All arithmetic runs in hand-written CUDA behind ``include/artspeech_b200.h``; there is no CPU
fallback (calls raise when the library or an sm_100 device is missing).
"""
__version__ = "0.1.0"
