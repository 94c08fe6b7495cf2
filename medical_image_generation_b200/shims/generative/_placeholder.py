def placeholder(name: str, why: str):
    """A class that imports fine and raises on construction."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(f"generative.{name} is outside the B200 hot path ({why}); install monai-generative "
                                  "to use it")

    return type(name.rsplit(".", 1)[-1], (), {"__init__": __init__, "__doc__": f"Placeholder for generative.{name}."})
