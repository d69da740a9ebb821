"""Stand-in for the chainer-chemistry pieces the hot-path files import (see ../README.md)."""
from . import config, functions, links  # noqa: F401
