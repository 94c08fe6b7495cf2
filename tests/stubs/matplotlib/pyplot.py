"""TEST SCAFFOLDING: every pyplot function is a no-op."""


def __getattr__(name):
    def _noop(*_a, **_k):
        return None
    return _noop
