"""No-op colors stand-in (see package docstring)."""


class ListedColormap:
    def __init__(self, colors=None, name=None, N=None):
        self.colors, self.name, self.N = colors, name, N


class Normalize:
    def __init__(self, *a, **k):
        pass
