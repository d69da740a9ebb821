from .connection.graph_linear import GraphLinear  # noqa: F401
from .connection.embed_atom_id import EmbedAtomID  # noqa: F401
from . import connection, embed_atom_id, update, readout  # noqa: F401
