"""Stand-in for `chainer` (see ../README.md): Link / Chain / ChainList bookkeeping + Variable on the oracle's NumPy tape."""
import contextlib

import numpy

from oracle import minichainer as _M

Variable = _M.Var


def as_variable(x):
    return x if isinstance(x, _M.Var) else _M.const(numpy.asarray(x))


def Parameter(initializer=None, shape=None, name=None):
    data = numpy.zeros(shape) if initializer is None or not isinstance(initializer, numpy.ndarray) else numpy.array(initializer)
    return _M.param(data)


class Link(object):
    def __init__(self):
        self.__dict__["_params"] = []
        self.__dict__["_children"] = []
        self.__dict__["_scope"] = False
        self.__dict__["xp"] = numpy

    @contextlib.contextmanager
    def init_scope(self):
        old = self._scope
        self.__dict__["_scope"] = True
        try:
            yield
        finally:
            self.__dict__["_scope"] = old

    def __setattr__(self, name, value):
        if self.__dict__.get("_scope"):
            if isinstance(value, Link) and name not in self._children:
                self._children.append(name)
            elif isinstance(value, _M.Var) and name not in self._params:
                self._params.append(name)
        object.__setattr__(self, name, value)

    def add_param(self, name, shape=None, dtype=numpy.float64, initializer=None):
        self._params.append(name)
        object.__setattr__(self, name, _M.param(numpy.zeros(shape, dtype)) if shape is not None else None)

    def register_persistent(self, name):
        pass

    def namedlinks(self, prefix=""):
        yield prefix or "/", self
        for c in self._children:
            for item in getattr(self, c).namedlinks(prefix + "/" + c):
                yield item

    def namedparams(self, prefix=""):
        for n in self._params:
            yield prefix + "/" + n, getattr(self, n)
        for c in self._children:
            for item in getattr(self, c).namedparams(prefix + "/" + c):
                yield item

    def params(self):
        for _, p in self.namedparams():
            yield p

    def cleargrads(self):
        for p in self.params():
            if p is not None:
                p.grad = None

    zerograds = cleargrads

    def to_gpu(self, device=None):
        return self

    def to_cpu(self):
        return self


class Chain(Link):
    pass


class ChainList(Link):
    def __init__(self, *links):
        Link.__init__(self)
        self.__dict__["_list"] = []
        for l in links:
            self.add_link(l)

    def add_link(self, link):
        name = str(len(self._list))
        self._list.append(link)
        self._children.append(name)
        object.__setattr__(self, name, link)

    def __getitem__(self, i):
        return self._list[i]

    def __iter__(self):
        return iter(self._list)

    def __len__(self):
        return len(self._list)


class Function(object):
    pass


class FunctionNode(object):
    pass


@contextlib.contextmanager
def no_backprop_mode():
    yield


@contextlib.contextmanager
def using_config(name, value):
    yield


class _Config(object):
    train = True
    enable_backprop = True


config = _Config()

from . import functions, links, cuda, backends, initializers, link, variable  # noqa: E402,F401
