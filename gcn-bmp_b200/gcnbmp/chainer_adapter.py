"""Chainer / CuPy side of the drop-in: `chainer.FunctionNode`s that hand CuPy device pointers to the same
C-ABI entry points the Torch host code uses, so the reference's own links can call the B200 kernels from a Chainer process.

Import-guarded: Chainer and CuPy are not installable in the build image (no network; the reference is Python 2.7 era), so
this module imports cleanly without them (`AVAILABLE` is False) and raises an ImportError with instructions when one of its
functions is used.  The argument filling is the code of gcnbmp/functional.py (exercised by the GPU test-suite through Torch)
with `arr.data.ptr` in place of `tensor.data_ptr()` and `cupy.cuda.get_current_stream().ptr` in place of the Torch stream;
the Chainer glue itself (retain_inputs / backward signatures) could not be executed here.

Usage on the reference side (models/models/ggnn.py:72-106 becomes one call):

    from gcnbmp import chainer_adapter as B
    class GGNN(chainer.Chain):
        ...
        def __call__(self, atom_array, adj, is_real_node=None):
            h0, hT = B.ggnn_encode(self, atom_array, adj)          # embed + all T GGNNUpdate steps, one launch
            self.atoms = hT
            return self.readout_layers[0](hT, h0, is_real_node)    # or B.readout(...)
"""
import ctypes as C

from . import _capi as K

try:                                    # pragma: no cover - neither package exists in the build image
    import chainer
    import cupy
    AVAILABLE = True
    _FunctionNode = chainer.FunctionNode
except Exception:                       # ImportError, or a broken CUDA runtime behind CuPy
    chainer = cupy = None
    AVAILABLE = False
    _FunctionNode = object


def _require():
    if not AVAILABLE:
        raise ImportError("gcnbmp.chainer_adapter needs `chainer` and `cupy` in the calling process (they are not part of this "
                          "repository's image); the Torch host code in gcnbmp.links / gcnbmp.functional is the tested path")


def _p(a):
    """cupy.ndarray (C-contiguous) -> device pointer for the ctypes structs of include/gcnbmp.h"""
    return None if a is None else C.c_void_p(a.data.ptr)


def _stream():
    return C.c_void_p(cupy.cuda.get_current_stream().ptr)


def _gru_arrays(gru):
    """chainer.links.GRU (= StatefulGRU) -> the 12 arrays in bmp_gru_t order"""
    out = []
    for name in ("W_r", "U_r", "W_z", "U_z", "W", "U"):
        lin = getattr(gru, name)
        out += [lin.W, lin.b]
    return out


def _fill_gru(dst, arrays):
    names = ("W_r", "b_Wr", "U_r", "b_Ur", "W_z", "b_Wz", "U_z", "b_Uz", "W", "b_W", "U", "b_U")
    for n, a in zip(names, arrays):
        setattr(dst, n, _p(a))


class GGNNEncode(_FunctionNode):
    """embed -> T x GGNNUpdate through bmp_ggnn_forward / bmp_ggnn_backward (fp32 mode: parity <= 1e-4 with the Chainer path).
    inputs: embed_W, then per distinct message layer (W, b), then per distinct GRU its 12 arrays -- all float32 cupy arrays;
    `plan` = [(message index, gru index, stateful)] per step as in gcnbmp.links (tied GGNN: [(0, 0, t > 0) for t in range(T)])."""

    def __init__(self, atoms, adj, plan, n_msg, n_gru):
        _require()
        self.atoms = cupy.ascontiguousarray(atoms, dtype=cupy.int32)
        self.adj = cupy.ascontiguousarray(adj, dtype=cupy.float32)
        self.plan, self.n_msg, self.n_gru = tuple(plan), n_msg, n_gru

    def _split(self, arrays):
        embed_W = arrays[0]
        msg = [(arrays[1 + 2 * i], arrays[2 + 2 * i]) for i in range(self.n_msg)]
        base = 1 + 2 * self.n_msg
        gru = [arrays[base + 12 * i: base + 12 * (i + 1)] for i in range(self.n_gru)]
        return embed_W, msg, gru

    def forward_gpu(self, inputs):
        self.retain_inputs(tuple(range(len(inputs))))
        embed_W, msg, gru = self._split(inputs)
        mb, E, N, _ = self.adj.shape
        H, T = msg[0][0].shape[1], len(self.plan)
        a = K.GgnnFwd()
        a.mb, a.n_atoms, a.hidden, a.n_edge, a.n_steps, a.mode = mb, N, H, E, T, K.MODE_F32
        a.atoms, a.embed_W, a.n_atom_types, a.adj = _p(self.atoms), _p(embed_W), embed_W.shape[0], _p(self.adj)
        for t, (mi, gi, st) in enumerate(self.plan):
            a.msg_W[t], a.msg_b[t] = _p(msg[mi][0]), _p(msg[mi][1])
            _fill_gru(a.gru[t], gru[gi])
            a.stateful[t] = int(st)
        rows = mb * N
        self.Hs = cupy.empty((T + 1, mb, N, H), dtype=cupy.float32)
        self.Ms = cupy.empty((T, rows, H), dtype=cupy.float32)
        self.Gs = cupy.empty((T, rows, 3 * H), dtype=cupy.float32)
        self.RSs = cupy.empty((T, rows, H), dtype=cupy.float32)
        a.Hs, a.Ms, a.Gs, a.RSs = _p(self.Hs), _p(self.Ms), _p(self.Gs), _p(self.RSs)
        K.check(K.lib.bmp_ggnn_forward(C.byref(a), _stream()))
        return self.Hs[0], self.Hs[T]                      # h_0 (for the readout) and h_T (get_atom_array)

    def backward(self, target_input_indexes, grad_outputs):
        params = [v.array for v in self.get_retained_inputs()]
        embed_W, msg, gru = self._split(params)
        mb, E, N, _ = self.adj.shape
        H, T = msg[0][0].shape[1], len(self.plan)
        rows = mb * N
        dHs = cupy.zeros((T + 1, mb, N, H), dtype=cupy.float32)
        if grad_outputs[0] is not None:
            dHs[0] = grad_outputs[0].array
        if grad_outputs[1] is not None:
            dHs[T] = grad_outputs[1].array
        grads = [cupy.zeros_like(p) for p in params]
        _, gmsg, ggru = self._split(grads)
        a = K.GgnnBwd()
        a.mb, a.n_atoms, a.hidden, a.n_edge, a.n_steps, a.mode = mb, N, H, E, T, K.MODE_F32
        a.adj = _p(self.adj)
        for t, (mi, gi, st) in enumerate(self.plan):
            a.msg_W[t] = _p(msg[mi][0])
            _fill_gru(a.gru[t], gru[gi])
            a.stateful[t] = int(st)
            a.d_msg_W[t], a.d_msg_b[t] = _p(gmsg[mi][0]), _p(gmsg[mi][1])
            _fill_gru(a.d_gru[t], ggru[gi])
        Ps = cupy.empty((T, rows, E * H), dtype=cupy.float32)
        a.Hs, a.Ms, a.RSs, a.Gs, a.Ps, a.dHs = _p(self.Hs), _p(self.Ms), _p(self.RSs), _p(self.Gs), _p(Ps), _p(dHs)
        K.check(K.lib.bmp_ggnn_backward(C.byref(a), _stream()))
        K.check(K.lib.bmp_embed_backward(_p(self.atoms), _p(dHs[0]), _p(grads[0]), rows, H, grads[0].shape[0], _stream()))
        return tuple(chainer.Variable(g) for g in grads)


def ggnn_encode(ggnn_link, atom_array, adj):
    """models/models/ggnn.py:72-106 (embed + the T-step loop) for the reference's modular GGNN chain -> (h_0, h_T) Variables."""
    _require()
    ups = list(ggnn_link.update_layers)
    tied = len(ups) == 1
    T = ggnn_link.n_layers
    plan = [(0, 0, t > 0) for t in range(T)] if tied else [(t, t, False) for t in range(T)]
    params = [ggnn_link.embed.W]
    for u in ups:
        params += [u.graph_linear.W, u.graph_linear.b]
    for u in ups:
        params += _gru_arrays(u.update_layer)
    return GGNNEncode(atom_array, adj, plan, len(ups), len(ups)).apply(tuple(params))
