"""Test-infrastructure shim for the `attrdict` package (reference test.py:12)."""


class AttrDict(dict):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.__dict__ = self
