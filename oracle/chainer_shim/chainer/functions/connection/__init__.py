from . import bilinear  # noqa: F401
