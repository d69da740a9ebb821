"""chainer.cuda / chainer.backends.cuda: CPU only."""
import numpy


class _NoCupy(object):
    ndarray = type("ndarray", (), {})


cupy = _NoCupy()
ndarray = _NoCupy.ndarray


def get_array_module(*args):
    return numpy


def to_gpu(x, device=None):
    return x


def to_cpu(x):
    return x


class _Dev(object):
    id = -1

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def get_device_from_array(*arrays):
    return _Dev()
