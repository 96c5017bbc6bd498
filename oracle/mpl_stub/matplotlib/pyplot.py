"""No-op pyplot stand-in (see package docstring)."""


class _Nop:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter(())


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Nop()
