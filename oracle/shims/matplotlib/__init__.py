"""Test-infrastructure shim: the reference imports matplotlib at import time only
(Vocoder/vocoder_utils.py:3-7, utils.py:5); nothing on the synthesis path plots."""


def use(*a, **kw):
    return None
