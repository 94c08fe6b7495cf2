"""TEST SCAFFOLDING: stand-in for matplotlib (absent from this image) so the unmodified medimgen trainers import.
Only `matplotlib.use` and `matplotlib.pyplot` are referenced at import time; plotting is never exercised by the tests."""


def use(*_a, **_k):
    return None
