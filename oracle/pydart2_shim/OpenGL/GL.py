"""Stub: lets the unmodified reference env modules import and 'open' a viewer without GL
(test infrastructure).  Every attribute is a no-op callable / zero constant."""


class _Noop(int):
    def __call__(self, *a, **kw):
        return 0


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Noop(0)
