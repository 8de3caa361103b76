from . import stereographic  # noqa: F401


class Sphere:  # referenced by name only
    pass
