"""No-op matplotlib stand-in so the reference's pipeline modules import (they pull in
matplotlib at module level only for live display).  Test infrastructure only."""
__version__ = "0.0-stub"


def use(*a, **k):
    pass


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return lambda *a, **k: None
