from . import graph_linear, embed_atom_id  # noqa: F401
