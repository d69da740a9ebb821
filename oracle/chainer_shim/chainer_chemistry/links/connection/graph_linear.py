"""chainer_chemistry GraphLinear: links.Linear applied to the last axis of a (s0, s1, in) array."""
from chainer import links
from oracle import minichainer as _M


class GraphLinear(links.Linear):
    def __call__(self, x):
        x = _M.as_var(x)
        s0, s1, s2 = x.shape
        y = links.Linear.__call__(self, _M.reshape(x, (s0 * s1, s2)))
        return _M.reshape(y, (s0, s1, self.out_size))
