"""No-op pyplot stand-in: every attribute is a callable that does nothing."""


def __getattr__(name):
    return lambda *a, **k: None
