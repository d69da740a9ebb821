"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Pin status: the op SEQUENCES are pinned to the reference's own link files, the Chainer
PRIMITIVES on this tape remain **parity unpinned** (see below).

A ~300-line reverse-mode tape over NumPy that restates the semantics of the
Chainer `functions` / `links` the GCN-BMP hot path executes (the reference
ships none of that arithmetic: it lives in Chainer v5/v6-era and a
chainer-chemistry v0.5-era fork, neither vendored nor pinned, neither
installable here -- see DESIGN.md).  The reference has no tests, golden
vectors or fixtures for this path, and Chainer cannot run
in this image, so this oracle is pinned only by: an independent Torch-autograd
twin (tests/torch_twin.py), finite differences, and algebraic identities.
That is why the header says "parity unpinned" for the primitives.  What IS pinned: oracle/chainer_shim exposes this tape
under the module names `chainer` / `chainer_chemistry`, so the reference's OWN hot-path files (/root/reference/models/*.py,
unmodified) execute in the build container; tests/golden/make_reference_golden.py generates tests/golden/ref_*.npz from
them (GGNN modular + monolithic files, GGNNUpdate, RelGCN, readout, the three fine co-attentions, HolE / MLP / SymMLP /
NTN / DistMult) and tests/test_reference_golden.py holds both the oracle (1e-10, fp64) and the CUDA links (1e-4) to them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product
(gcn-bmp_b200/) never does.

Semantics restated (published Chainer behaviour):
  * links.Linear            y = x W^T + b, W is (out, in)
  * links.EmbedID           y = W[ids]
  * links.Bilinear          y = sum_ij e1_i W_ijk e2_j + e1 V1 + e2 V2 + b
  * links.GRU (= StatefulGRU): first call after reset_state has no U terms and
    returns z*h_bar; later calls return z*h_bar + (1-z)*h
  * functions.softmax       default axis=1
  * functions.fft / ifft    over the last axis, on (real, imag) pairs; ifft
                            scaled by 1/D
  * functions.sigmoid_cross_entropy  mean over non-ignored (-1) elements of
    -(x (t - [x>=0]) - log1p(exp(-|x|)))
  * chainer_chemistry GraphLinear = Linear over the last axis of a 3-D array
"""
import numpy as np


class Var(object):
    """A node on the tape: value, accumulated gradient, and how to push it."""
    __slots__ = ("data", "grad", "_src", "_push", "requires")

    def __init__(self, data, src=(), push=None, requires=None):
        self.data = data
        self.grad = None
        self._src = src
        self._push = push
        if requires is None:
            requires = any(s.requires for s in src)
        self.requires = requires

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return self.data.ndim

    def __add__(self, o):
        return add(self, o)

    def __mul__(self, o):
        return mul(self, o)

    def __sub__(self, o):
        return sub(self, o)

    def __rsub__(self, o):
        return sub(o, self)

    # conveniences the reference's own files rely on when they run over oracle/chainer_shim (chainer.Variable has them too)
    __radd__ = __add__
    __rmul__ = __mul__
    __array_ufunc__ = None        # ndarray (op) Var defers to Var's reflected operator, as with chainer.Variable

    def __len__(self):
        return len(self.data)

    def __truediv__(self, o):
        return div(self, o)

    def __rtruediv__(self, o):
        return div(o, self)

    def __neg__(self):
        return sub(0.0, self)

    def __getitem__(self, idx):
        return getitem(self, idx)

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    @property
    def array(self):
        return self.data

    def backward(self, seed=None):
        order, seen = [], set()

        def visit(v):
            stack = [(v, iter(v._src))]
            seen.add(id(v))
            while stack:
                node, it = stack[-1]
                nxt = next(it, None)
                if nxt is None:
                    order.append(node)
                    stack.pop()
                elif id(nxt) not in seen:
                    seen.add(id(nxt))
                    stack.append((nxt, iter(nxt._src)))
        visit(self)
        self.grad = np.ones_like(self.data) if seed is None else np.asarray(seed, self.data.dtype)
        for node in reversed(order):
            if node._push is None or node.grad is None:
                continue
            for s, g in zip(node._src, node._push(node.grad)):
                if g is None or not s.requires:
                    continue
                s.grad = g if s.grad is None else s.grad + g


def param(a):
    return Var(np.asarray(a), requires=True)


def const(a):
    return Var(np.asarray(a), requires=False)


def as_var(a):
    return a if isinstance(a, Var) else const(a)


def _unbroadcast(g, shape):
    while g.ndim > len(shape):
        g = g.sum(axis=0)
    for ax, n in enumerate(shape):
        if n == 1 and g.shape[ax] != 1:
            g = g.sum(axis=ax, keepdims=True)
    return g


# ---- element-wise -----------------------------------------------------------
def add(a, b):
    a, b = as_var(a), as_var(b)
    return Var(a.data + b.data, (a, b),
               lambda g: (_unbroadcast(g, a.shape), _unbroadcast(g, b.shape)))


def sub(a, b):
    a, b = as_var(a), as_var(b)
    if not isinstance(a.data, np.ndarray) or a.data.ndim == 0:
        a = const(np.asarray(a.data, dtype=b.dtype))
    return Var(a.data - b.data, (a, b),
               lambda g: (_unbroadcast(g, a.shape), _unbroadcast(-g, b.shape)))


def mul(a, b):
    a, b = as_var(a), as_var(b)
    return Var(a.data * b.data, (a, b),
               lambda g: (_unbroadcast(g * b.data, a.shape), _unbroadcast(g * a.data, b.shape)))


def div(a, b):
    a, b = as_var(a), as_var(b)
    if not isinstance(a.data, np.ndarray) or a.data.ndim == 0:
        a = const(np.asarray(a.data, dtype=b.dtype))
    return Var(a.data / b.data, (a, b),
               lambda g: (_unbroadcast(g / b.data, a.shape), _unbroadcast(-g * a.data / (b.data * b.data), b.shape)))


def where(cond, x, y):
    """functions.where(condition (plain bool array), x, y)."""
    x, y = as_var(x), as_var(y)
    c = np.asarray(cond)
    return Var(np.where(c, x.data, y.data), (x, y),
               lambda g: (_unbroadcast(np.where(c, g, 0), x.shape), _unbroadcast(np.where(c, 0, g), y.shape)))


def getitem(x, idx):
    """Variable.__getitem__ (basic slicing / None-insertion as the reference uses it)."""
    def push(g):
        gx = np.zeros_like(x.data)
        np.add.at(gx, idx, g)
        return (gx,)
    return Var(x.data[idx], (x,), push)


def sigmoid(x):
    y = (0.5 * (1.0 + np.tanh(0.5 * x.data))).astype(x.dtype, copy=False)   # overflow-free logistic
    return Var(y, (x,), lambda g: (g * y * (1 - y),))


def tanh(x):
    y = np.tanh(x.data)
    return Var(y, (x,), lambda g: (g * (1 - y * y),))


def relu(x):
    m = x.data > 0
    return Var(x.data * m, (x,), lambda g: (g * m,))


def identity(x):
    return x


def copy(x):
    return Var(x.data.copy(), (x,), lambda g: (g,))


def linear_interpolate(p, x, y):
    """chainer.functions.linear_interpolate: p*x + (1-p)*y."""
    return add(mul(p, x), mul(sub(np.asarray(1, p.dtype), p), y))


# ---- shape ops --------------------------------------------------------------
def reshape(x, shape):
    return Var(x.data.reshape(shape), (x,), lambda g: (g.reshape(x.shape),))


def transpose(x, axes):
    inv = np.argsort(axes)
    return Var(np.transpose(x.data, axes), (x,), lambda g: (np.transpose(g, inv),))


def expand_dims(x, axis):
    return Var(np.expand_dims(x.data, axis), (x,), lambda g: (g.reshape(x.shape),))


def tile(x, reps):
    """functions.tile -- materialises the copies, exactly as Chainer does."""
    reps = tuple(reps)
    assert len(reps) == x.ndim

    def push(g):
        shp = []
        for r, n in zip(reps, x.shape):
            shp += [r, n]
        return (g.reshape(shp).sum(axis=tuple(range(0, 2 * x.ndim, 2))),)
    return Var(np.tile(x.data, reps), (x,), push)


def concat(xs, axis=1):
    sizes = np.cumsum([v.shape[axis] for v in xs])[:-1]
    return Var(np.concatenate([v.data for v in xs], axis=axis), tuple(xs),
               lambda g: tuple(np.split(g, sizes, axis=axis)))


def sum_(x, axis=None):
    def push(g):
        if axis is None:
            return (np.broadcast_to(g, x.shape).astype(x.dtype),)
        return (np.broadcast_to(np.expand_dims(g, axis), x.shape).astype(x.dtype),)
    return Var(x.data.sum(axis=axis), (x,), push)


def mean(x, axis):
    n = x.shape[axis]
    return Var(x.data.mean(axis=axis), (x,),
               lambda g: ((np.broadcast_to(np.expand_dims(g, axis), x.shape) / n).astype(x.dtype),))


def stack(xs, axis=0):
    """functions.stack."""
    xs = tuple(as_var(x) for x in xs)
    return Var(np.stack([v.data for v in xs], axis=axis), xs,
               lambda g: tuple(np.take(g, i, axis=axis) for i in range(len(xs))))


def max_(x, axis=None, keepdims=False):
    """functions.max: the gradient goes to EVERY position that equals the maximum (chainer's `cond = x == y`), undivided."""
    y = x.data.max(axis=axis, keepdims=True)
    cond = (x.data == y)
    out = y if keepdims else np.squeeze(y, axis=axis)

    def push(g):
        gk = g if keepdims else np.expand_dims(g, axis)
        return ((cond * gk).astype(x.dtype),)
    return Var(out, (x,), push)


def maximum(a, b):
    """functions.maximum: gradient to `a` where a >= b, to `b` elsewhere."""
    a, b = as_var(a), as_var(b)
    cond = a.data >= b.data
    return Var(np.maximum(a.data, b.data), (a, b),
               lambda g: (_unbroadcast(np.where(cond, g, 0), a.shape), _unbroadcast(np.where(cond, 0, g), b.shape)))


def normalize(x, eps=1e-5, axis=1):
    """functions.normalize: x / (||x||_2 + eps) along `axis` (the epsilon is added to the norm, chainer v2+)."""
    n = np.sqrt((x.data * x.data).sum(axis=axis, keepdims=True))
    s = 1.0 / (n + eps)
    y = x.data * s

    def push(g):
        dot = (x.data * g).sum(axis=axis, keepdims=True)
        with np.errstate(divide="ignore", invalid="ignore"):
            corr = np.where(n > 0, x.data * dot * s * s / n, 0.0)
        return ((g * s - corr).astype(x.dtype),)
    return Var(y.astype(x.dtype, copy=False), (x,), push)


# ---- contractions -----------------------------------------------------------
def matmul(a, b):
    """functions.matmul on >=2-D operands (batched over leading axes)."""
    a, b = as_var(a), as_var(b)

    def push(g):
        ga = np.matmul(g, np.swapaxes(b.data, -1, -2))
        gb = np.matmul(np.swapaxes(a.data, -1, -2), g)
        return (_unbroadcast(ga, a.shape), _unbroadcast(gb, b.shape))
    return Var(np.matmul(a.data, b.data), (a, b), push)


def linear(x, W, b=None):
    """links.Linear on a 2-D input: x W^T + b."""
    y = x.data @ W.data.T
    if b is not None:
        y = y + b.data
    src = (x, W) if b is None else (x, W, b)

    def push(g):
        out = (g @ W.data, g.T @ x.data)
        return out if b is None else out + (g.sum(axis=0),)
    return Var(y, src, push)


def graph_linear(x, W, b=None):
    """chainer_chemistry GraphLinear: Linear over the last axis of (s0, s1, in)."""
    s0, s1, s2 = x.shape
    y = linear(reshape(x, (s0 * s1, s2)), W, b)
    return reshape(y, (s0, s1, W.shape[0]))


def embed_id(ids, W):
    ids = np.asarray(ids)

    def push(g):
        gw = np.zeros_like(W.data)
        np.add.at(gw, ids.reshape(-1), g.reshape(-1, W.shape[1]))
        return (gw,)
    return Var(W.data[ids], (W,), push)


def bilinear(e1, e2, W, V1=None, V2=None, b=None):
    """links.Bilinear: y_l = sum_jk e1_j W_jkl e2_k + e1 V1 + e2 V2 + b.
    Chainer materialises the (B, J, K) outer product e1 (x) e2 and tensordots it with W
    (8.6 GB at B = 32*64*64, J = K = 128); the contraction here goes through BLAS GEMMs
    instead (same result, far cheaper), which makes this CPU baseline optimistic for the
    reference."""
    L = W.shape[2]
    t = [e1.data @ W.data[:, :, l] for l in range(L)]                 # (B, K) each
    y = np.stack([(t[l] * e2.data).sum(axis=1) for l in range(L)], axis=1)
    src = [e1, e2, W]
    if V1 is not None:
        y = y + e1.data @ V1.data + e2.data @ V2.data + b.data
        src += [V1, V2, b]

    def push(g):
        ge1 = sum((g[:, l:l + 1] * e2.data) @ W.data[:, :, l].T for l in range(L))
        ge2 = sum(g[:, l:l + 1] * t[l] for l in range(L))
        gW = np.stack([(e1.data * g[:, l:l + 1]).T @ e2.data for l in range(L)], axis=2)
        out = [ge1, ge2, gW]
        if V1 is not None:
            out[0] = out[0] + g @ V1.data.T
            out[1] = out[1] + g @ V2.data.T
            out += [e1.data.T @ g, e2.data.T @ g, g.sum(axis=0)]
        return tuple(out)
    return Var(y.astype(e1.dtype, copy=False), tuple(src), push)


# ---- normalisers / transforms ----------------------------------------------
def softmax(x, axis=1):
    z = x.data - x.data.max(axis=axis, keepdims=True)
    e = np.exp(z)
    y = e / e.sum(axis=axis, keepdims=True)

    def push(g):
        return (y * (g - (g * y).sum(axis=axis, keepdims=True)),)
    return Var(y, (x,), push)


def fft(pair):
    re, im = pair
    z = np.fft.fft(re.data.astype(np.complex128) + 1j * im.data, axis=-1)
    dt = re.dtype
    out_re = Var(z.real.astype(dt), (re, im), None)
    out_im = Var(z.imag.astype(dt), (re, im), None)

    # adjoint of the unnormalised DFT is the unnormalised inverse-sign DFT
    def push_re(g):
        w = np.fft.ifft(g.astype(np.complex128), axis=-1) * g.shape[-1]
        return (w.real.astype(dt), w.imag.astype(dt))

    def push_im(g):
        w = np.fft.ifft(1j * g.astype(np.complex128), axis=-1) * g.shape[-1]
        return (w.real.astype(dt), w.imag.astype(dt))
    out_re._push, out_im._push = push_re, push_im
    return out_re, out_im


def ifft(pair):
    re, im = pair
    z = np.fft.ifft(re.data.astype(np.complex128) + 1j * im.data, axis=-1)
    dt = re.dtype
    out_re = Var(z.real.astype(dt), (re, im), None)
    out_im = Var(z.imag.astype(dt), (re, im), None)

    def push_re(g):
        w = np.fft.fft(g.astype(np.complex128), axis=-1) / g.shape[-1]
        return (w.real.astype(dt), w.imag.astype(dt))

    def push_im(g):
        w = np.fft.fft(1j * g.astype(np.complex128), axis=-1) / g.shape[-1]
        return (w.real.astype(dt), w.imag.astype(dt))
    out_re._push, out_im._push = push_re, push_im
    return out_re, out_im


def sigmoid_cross_entropy(x, t):
    """functions.sigmoid_cross_entropy(normalize=True, reduce='mean')."""
    t = np.asarray(t)
    keep = (t != -1)
    cnt = max(int(keep.sum()), 1)
    xd = x.data
    per = -(keep * (xd * (t - (xd >= 0)) - np.log1p(np.exp(-np.abs(xd)))))
    val = np.asarray(per.sum() / cnt, dtype=x.dtype)

    def push(g):
        y = 0.5 * (1.0 + np.tanh(0.5 * xd))
        return ((g * keep * (y - t) / cnt).astype(x.dtype),)
    return Var(val, (x,), push)
