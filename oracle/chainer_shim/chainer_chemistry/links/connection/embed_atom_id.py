"""chainer_chemistry EmbedAtomID: links.EmbedID over atom ids (int32), any leading shape."""
from chainer import links


class EmbedAtomID(links.EmbedID):
    pass
