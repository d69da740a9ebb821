"""Stand-in for `chainer.functions`: the functions the reference's hot-path files call, on the oracle's NumPy tape."""
import numpy

from oracle import minichainer as _M

reshape, transpose, expand_dims, tile, concat = _M.reshape, _M.transpose, _M.expand_dims, _M.tile, _M.concat
softmax, matmul, add, sigmoid, tanh, relu, identity = _M.softmax, _M.matmul, _M.add, _M.sigmoid, _M.tanh, _M.relu, _M.identity
fft, ifft, linear, bilinear, sigmoid_cross_entropy = _M.fft, _M.ifft, _M.linear, _M.bilinear, _M.sigmoid_cross_entropy
linear_interpolate, embed_id, where = _M.linear_interpolate, _M.embed_id, _M.where


def sum(x, axis=None):  # noqa: A001 (chainer's name)
    return _M.sum_(_M.as_var(x), axis=axis)


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
