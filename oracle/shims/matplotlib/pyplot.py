rcParams = {}


def switch_backend(*a, **k):
    pass


def __getattr__(name):
    def _noop(*a, **k):
        raise RuntimeError("matplotlib stub: plotting is out of scope (%s)" % name)

    return _noop
