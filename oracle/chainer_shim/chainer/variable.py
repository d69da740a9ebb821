"""chainer.variable: Parameter(initializer, shape) as models/mlp.py:166-167 builds it."""
import numpy

from oracle import minichainer as _M

Variable = _M.Var


def Parameter(initializer=None, shape=None, name=None):
    return _M.param(numpy.zeros(shape))
