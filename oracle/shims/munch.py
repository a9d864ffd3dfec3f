"""Test-infrastructure shim: minimal stand-in for the `munch` package (absent in this image).

Only used by oracle/ref_loader.py so that /root/reference/models.py (models.py:6) imports.
"""


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
