"""Stand-in for `chainer.functions`: the functions the reference's hot-path files call, on the oracle's NumPy tape."""
import numpy

from oracle import minichainer as _M



def _v(fn, n=1):
    """chainer functions take raw arrays as well as Variables: wrap the first n positional arguments"""
    def wrapped(*args, **kw):
        args = tuple(_M.as_var(a) if i < n and isinstance(a, numpy.ndarray) else a for i, a in enumerate(args))
        return fn(*args, **kw)
    wrapped.__name__ = fn.__name__
    return wrapped


reshape, transpose, expand_dims, tile = _v(_M.reshape), _v(_M.transpose), _v(_M.expand_dims), _v(_M.tile)
softmax, sigmoid, tanh, relu, identity = _v(_M.softmax), _v(_M.sigmoid), _v(_M.tanh), _v(_M.relu), _v(_M.identity)
matmul, add = _M.matmul, _M.add
fft, ifft, linear, bilinear, sigmoid_cross_entropy = _M.fft, _M.ifft, _M.linear, _M.bilinear, _M.sigmoid_cross_entropy
linear_interpolate, embed_id, where = _M.linear_interpolate, _M.embed_id, _M.where


def stack(xs, axis=0):
    return _M.stack(xs, axis=axis)


def max(x, axis=None, keepdims=False):  # noqa: A001 (chainer's name)
    return _M.max_(_M.as_var(x), axis=axis, keepdims=keepdims)


maximum = _M.maximum


def normalize(x, eps=1e-5, axis=1):
    return _M.normalize(_M.as_var(x), eps=eps, axis=axis)


def concat(xs, axis=1):
    return _M.concat(tuple(_M.as_var(x) for x in xs), axis=axis)


def sum(x, axis=None, keepdims=False):  # noqa: A001 (chainer's name)
    y = _M.sum_(_M.as_var(x), axis=axis)
    return _M.expand_dims(y, axis) if keepdims and axis is not None else y


def mean(x, axis=None):
    return _M.mean(_M.as_var(x), axis)


def copy(x, dst=None):
    return _M.copy(_M.as_var(x))


def broadcast_to(x, shape):
    x = _M.as_var(x)
    return _M.Var(numpy.broadcast_to(x.data, shape).copy(), (x,), lambda g: (_M._unbroadcast(g, x.shape),))


def squeeze(x, axis=None):
    x = _M.as_var(x)
    return _M.reshape(x, numpy.squeeze(x.data, axis=axis).shape)


def dropout(x, ratio=0.5):
    if ratio != 0.0:
        raise NotImplementedError("chainer_shim: dropout with ratio > 0 (stochastic) is outside the parity path")
    return x
