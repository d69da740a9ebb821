"""Stand-in for `chainer.links`: Linear, GRU (= StatefulGRU), Bilinear, EmbedID -- Chainer's published forward formulas."""
import numpy

from oracle import minichainer as _M
from . import Link, Chain


class Linear(Link):
    """y = x W^T + b; W (out, in); in_size None = taken from the first input; inputs with ndim > 2 are flattened to 2-D."""

    def __init__(self, in_size, out_size=None, nobias=False, initialW=None, initial_bias=None):
        Link.__init__(self)
        if out_size is None:
            in_size, out_size = None, in_size
        self.__dict__.update(in_size=in_size, out_size=out_size, nobias=nobias)
        with self.init_scope():
            self.W = _M.param(numpy.zeros((out_size, in_size))) if in_size is not None else None
            if "W" not in self._params:
                self._params.append("W")
            self.b = None if nobias else _M.param(numpy.zeros((out_size,)))
            if not nobias and "b" not in self._params:
                self._params.append("b")

    def __call__(self, x):
        x = _M.as_var(x)
        if self.W is None:
            raise RuntimeError("chainer_shim: lazily-shaped Linear called before its W was loaded")
        if x.ndim > 2:
            x = _M.reshape(x, (x.shape[0], -1))
        return _M.linear(x, self.W, self.b)


class StatefulGRU(Chain):
    """chainer.links.GRU: first call after reset_state() has no state (out = z * h_bar), later calls are the full GRU."""

    def __init__(self, in_size, out_size):
        Chain.__init__(self)
        with self.init_scope():
            self.W_r, self.U_r = Linear(in_size, out_size), Linear(out_size, out_size)
            self.W_z, self.U_z = Linear(in_size, out_size), Linear(out_size, out_size)
            self.W, self.U = Linear(in_size, out_size), Linear(out_size, out_size)
        self.__dict__["h"] = None

    def reset_state(self):
        self.__dict__["h"] = None

    def set_state(self, h):
        self.__dict__["h"] = _M.as_var(h)

    def __call__(self, x):
        z, h_bar = self.W_z(x), self.W(x)
        if self.h is not None:
            r = _M.sigmoid(_M.add(self.W_r(x), self.U_r(self.h)))
            z = _M.add(z, self.U_z(self.h))
            h_bar = _M.add(h_bar, self.U(_M.mul(r, self.h)))
        z, h_bar = _M.sigmoid(z), _M.tanh(h_bar)
        h_new = _M.linear_interpolate(z, h_bar, self.h) if self.h is not None else _M.mul(z, h_bar)
        self.__dict__["h"] = h_new
        return h_new


GRU = StatefulGRU


class Bilinear(Link):
    def __init__(self, left_size, right_size, out_size, nobias=False, initialW=None, initial_bias=None):
        Link.__init__(self)
        self.__dict__["nobias"] = nobias
        with self.init_scope():
            self.W = _M.param(numpy.zeros((left_size, right_size, out_size)))
            if not nobias:
                self.V1 = _M.param(numpy.zeros((left_size, out_size)))
                self.V2 = _M.param(numpy.zeros((right_size, out_size)))
                self.b = _M.param(numpy.zeros((out_size,)))

    def __call__(self, e1, e2):
        if self.nobias:
            return _M.bilinear(_M.as_var(e1), _M.as_var(e2), self.W)
        return _M.bilinear(_M.as_var(e1), _M.as_var(e2), self.W, self.V1, self.V2, self.b)


class EmbedID(Link):
    def __init__(self, in_size, out_size, initialW=None, ignore_label=None):
        Link.__init__(self)
        with self.init_scope():
            self.W = _M.param(numpy.zeros((in_size, out_size)))

    def __call__(self, x):
        return _M.embed_id(numpy.asarray(getattr(x, "data", x)), self.W)
