"""No-op pyplot stand-in: every attribute is a callable that swallows its arguments and returns an object on which any further
attribute access, call, indexing or unpacking is again a no-op (the reference's M-models plot unconditionally:
`figure, axis = plt.subplots(n); axis[i].plot(...)`, `PyREMOT/library/plot.py:137`)."""


class _Nothing:
    def __call__(self, *a, **k):
        return _Nothing()

    def __getattr__(self, name):
        return _Nothing()

    def __getitem__(self, k):
        return _Nothing()

    def __iter__(self):
        return iter((_Nothing(), _Nothing()))


def __getattr__(name):
    return _Nothing()
