"""chainer.functions.connection.bilinear (models/mlp.py:15,176)."""
from oracle import minichainer as _M


def bilinear(e1, e2, W, V1=None, V2=None, b=None):
    return _M.bilinear(_M.as_var(e1), _M.as_var(e2), _M.as_var(W), V1, V2, b)
