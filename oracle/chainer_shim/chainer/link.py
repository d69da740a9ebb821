from . import Link, Chain, ChainList  # noqa: F401
