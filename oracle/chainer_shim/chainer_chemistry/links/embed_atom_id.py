from .connection.embed_atom_id import EmbedAtomID  # noqa: F401
