"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Op sequences pinned to fixtures generated from the reference's own files
(tests/golden/ref_*.npz); the Chainer primitives underneath remain **parity unpinned** (see minichainer.py).

CPU restatement of the GCN-BMP message-passing hot path, executing the SAME op
sequence the reference's Chainer links execute (same reshapes, transposes,
per-edge-type batched matmul, six separate GRU linears, tile-materialised
bilinear operands, FFT-based HolE), on the NumPy tape in minichainer.py.
Every function cites the reference file:line it follows (paths are relative to
the reference repository root).  Parameters are passed as a flat dict keyed by
Chainer-style parameter paths (what `Link.namedparams()` / an npz snapshot
would contain), e.g. ``update_layers/0/graph_linear/W``.

dtype: whatever the parameter/input arrays are (float32 = "reference CPU path",
float64 = ground truth for tolerance tests).
"""
import numpy as np

from . import minichainer as F

MAX_ATOMIC_NUM = 117  # chainer_chemistry.config.MAX_ATOMIC_NUM
ACT = {"identity": F.identity, "tanh": F.tanh, "relu": F.relu, "sigmoid": F.sigmoid}


class P(object):
    """View of a flat {path: Var} dict under a prefix."""

    def __init__(self, table, prefix=""):
        self.table, self.prefix = table, prefix

    def __getitem__(self, name):
        return self.table[self.prefix + name]

    def sub(self, name):
        return P(self.table, self.prefix + name + "/")

    def has(self, name):
        return (self.prefix + name) in self.table


def wrap_params(arrays, dtype=None):
    """{path: ndarray} -> {path: Var(requires grad)}."""
    return {k: F.param(np.asarray(v, dtype=dtype or v.dtype)) for k, v in arrays.items()}


def grads_of(table):
    return {k: (v.grad if v.grad is not None else np.zeros_like(v.data)) for k, v in table.items()}


# ----------------------------------------------------------------------------
# Chainer links.GRU (= StatefulGRU); used at models/update/ggnn_update.py:28,60
# ----------------------------------------------------------------------------
class StatefulGRU(object):
    def __init__(self, p):
        self.p, self.h = p, None

    def reset_state(self):
        self.h = None

    def _lin(self, name, x):
        q = self.p.sub(name)
        return F.linear(x, q["W"], q["b"])

    def __call__(self, x):
        z = self._lin("W_z", x)
        h_bar = self._lin("W", x)
        if self.h is not None:
            r = F.sigmoid(F.add(self._lin("W_r", x), self._lin("U_r", self.h)))
            z = F.add(z, self._lin("U_z", self.h))
            h_bar = F.add(h_bar, self._lin("U", F.mul(r, self.h)))
        z = F.sigmoid(z)
        h_bar = F.tanh(h_bar)
        if self.h is not None:
            out = F.linear_interpolate(z, h_bar, self.h)
        else:
            out = F.mul(z, h_bar)
        self.h = out
        return out


def _edge_messages(lin_out, mb, n, ch, n_edge):
    """(mb, n, ch*E) -> (mb, E, n, ch): channel-major / edge-minor split.
    models/update/ggnn_update.py:35-39, models/update/relgcn_update.py:33-37."""
    m = F.reshape(lin_out, (mb, n, ch, n_edge))
    return F.transpose(m, (0, 3, 1, 2))


# ----------------------------------------------------------------------------
# models/update/ggnn_update.py:31-63
# ----------------------------------------------------------------------------
class GGNNUpdate(object):
    def __init__(self, p, hidden_dim=16, num_edge_type=4, gru=None):
        self.p, self.E = p, num_edge_type
        self.update_layer = gru if gru is not None else StatefulGRU(p.sub("update_layer"))

    def message(self, h, adj):
        mb, n, ch = h.shape
        gl = self.p.sub("graph_linear")
        m = _edge_messages(F.graph_linear(h, gl["W"], gl["b"]), mb, n, ch, self.E)   # :34-39
        a = F.reshape(F.as_var(adj), (mb * self.E, n, n))                              # :41
        m = F.reshape(m, (mb * self.E, n, ch))                                         # :43
        m = F.matmul(a, m)                                                             # :45
        m = F.reshape(m, (mb, self.E, n, ch))                                          # :48
        return F.sum_(m, axis=1)                                                       # :49

    def __call__(self, h, adj):
        mb, n, ch = h.shape
        m = self.message(h, adj)
        x = F.concat((F.reshape(h, (mb * n, ch)), F.reshape(m, (mb * n, ch))), axis=1)  # :54-60
        return F.reshape(self.update_layer(x), (mb, n, ch))                             # :62

    def reset_state(self):
        self.update_layer.reset_state()


# ----------------------------------------------------------------------------
# models/update/relgcn_update.py:24-44
# ----------------------------------------------------------------------------
class RelGCNUpdate(object):
    def __init__(self, p, in_channels, out_channels, num_edge_type=4):
        self.p, self.E, self.cout = p, num_edge_type, out_channels

    def __call__(self, h, adj):
        mb, n, _ = h.shape
        ps, pe = self.p.sub("graph_linear_self"), self.p.sub("graph_linear_edge")
        hs = F.graph_linear(h, ps["W"], ps["b"])                                        # :28
        m = _edge_messages(F.graph_linear(h, pe["W"], pe["b"]), mb, n, self.cout, self.E)  # :32-37
        m = F.matmul(F.as_var(adj), m)                                                  # :40 (4-D batched)
        return F.add(hs, F.sum_(m, axis=1))                                             # :43-44


def rescale_adj(adj):
    """models/relgcn.py:20-28 -- column-degree normalisation over (edge, row)."""
    deg = adj.sum(axis=(1, 2))
    inv = 1.0 / np.where(deg != 0, deg, np.ones_like(deg))
    return (adj * inv[:, None, None, :]).astype(adj.dtype)


# ----------------------------------------------------------------------------
# models/readout/ggnn_readout.py:42-58   (variant R1)
# ----------------------------------------------------------------------------
class GGNNReadout(object):
    def __init__(self, p, out_dim, hidden_dim=16, nobias=False,
                 activation="identity", activation_agg="identity"):
        self.p, self.nobias = p, nobias
        self.act, self.act_agg = ACT[activation], ACT[activation_agg]

    def _gl(self, name, x):
        q = self.p.sub(name)
        return F.graph_linear(x, q["W"], None if self.nobias else q["b"])

    def __call__(self, h, h0=None, is_real_node=None):
        h1 = F.concat((h, h0), axis=2) if h0 is not None else h                        # :45
        g = F.mul(F.sigmoid(self._gl("i_layer", h1)), self.act(self._gl("j_layer", h1)))  # :47-49
        if is_real_node is not None:                                                    # :50-54
            mask = np.broadcast_to(np.asarray(is_real_node)[:, :, None], g.shape).astype(g.dtype)
            g = F.mul(g, F.const(mask))
        return self.act_agg(F.sum_(g, axis=1))                                          # :56


# ----------------------------------------------------------------------------
# models/models/ggnn.py:46-109  (modular GGNN: GGNNUpdate + GGNNReadout)
# ----------------------------------------------------------------------------
class GGNN(object):
    def __init__(self, p, out_dim, hidden_dim=16, n_layers=4, n_atom_types=MAX_ATOMIC_NUM,
                 concat_hidden=False, weight_tying=True, activation="identity", num_edge_type=4):
        self.p = p
        self.n_layers, self.concat_hidden, self.weight_tying = n_layers, concat_hidden, weight_tying
        n_msg = 1 if weight_tying else n_layers
        n_ro = n_layers if concat_hidden else 1
        self.update_layers = [GGNNUpdate(p.sub("update_layers/%d" % i), hidden_dim, num_edge_type)
                              for i in range(n_msg)]
        self.readout_layers = [GGNNReadout(p.sub("readout_layers/%d" % i), out_dim, hidden_dim,
                                           activation=activation, activation_agg=activation)
                               for i in range(n_ro)]
        self.atoms = None

    def reset_state(self):
        for u in self.update_layers:
            u.reset_state()

    def __call__(self, atom_array, adj, is_real_node=None):
        self.reset_state()                                                             # :88
        if np.asarray(getattr(atom_array, "data", atom_array)).ndim <= 2:              # :89
            h = F.embed_id(atom_array, self.p["embed/W"])
        else:
            h = F.as_var(atom_array)
        h0 = F.copy(h)                                                                 # :93
        gs = []
        for step in range(self.n_layers):                                              # :95
            u = self.update_layers[0 if self.weight_tying else step]
            h = u(h, adj)
            if self.concat_hidden:
                gs.append(self.readout_layers[step](h, h0, is_real_node))
        self.atoms = h
        if self.concat_hidden:
            return F.concat(gs, axis=1)                                                # :103
        return self.readout_layers[0](h, h0, is_real_node)                             # :105

    def get_atom_array(self):
        return self.atoms


# ----------------------------------------------------------------------------
# models/ggnn_att.py:220-268,338-346,589-664 (default flags) and
# models/ggnn_dev.py:69-122,135-172: monolithic GGNN -- per-step message
# GraphLinears, ONE shared stateful GRU, readout R2 (j sees only h), optional
# sum readout (ggnn_dev.py:165-167), final atom states exposed.
# ----------------------------------------------------------------------------
class GGNNMono(object):
    NUM_EDGE_TYPE = 4

    def __init__(self, p, out_dim, hidden_dim=16, n_layers=4, n_atom_types=MAX_ATOMIC_NUM,
                 concat_hidden=False, weight_tying=True, sum_readout=False):
        self.p, self.n_layers = p, n_layers
        self.concat_hidden, self.weight_tying, self.sum_readout = concat_hidden, weight_tying, sum_readout
        self.gru = StatefulGRU(p.sub("update_layer"))
        self.atoms_list = []

    def update(self, h, adj, step):
        mb, n, ch = h.shape
        idx = 0 if self.weight_tying else step
        q = self.p.sub("message_layers/%d" % idx)
        m = _edge_messages(F.graph_linear(h, q["W"], q["b"]), mb, n, ch, self.NUM_EDGE_TYPE)
        a = F.reshape(F.as_var(adj), (mb * self.NUM_EDGE_TYPE, n, n))
        m = F.matmul(a, F.reshape(m, (mb * self.NUM_EDGE_TYPE, n, ch)))
        m = F.sum_(F.reshape(m, (mb, self.NUM_EDGE_TYPE, n, ch)), axis=1)
        x = F.concat((F.reshape(h, (mb * n, ch)), F.reshape(m, (mb * n, ch))), axis=1)
        return F.reshape(self.gru(x), (mb, n, ch))

    def readout(self, h, h0, step=0):
        idx = step if self.concat_hidden else 0
        qi, qj = self.p.sub("i_layers/%d" % idx), self.p.sub("j_layers/%d" % idx)
        gate = F.sigmoid(F.graph_linear(F.concat((h, h0), axis=2), qi["W"], qi["b"]))
        g = F.mul(gate, F.graph_linear(h, qj["W"], qj["b"]))
        return F.sum_(g, axis=1)

    def __call__(self, atom_array, adj):
        self.gru.reset_state()
        self.atoms_list = []
        a = np.asarray(getattr(atom_array, "data", atom_array))
        h = F.embed_id(a, self.p["embed/W"]) if a.dtype == np.int32 else F.as_var(atom_array)
        h0 = F.copy(h)
        gs = []
        for step in range(self.n_layers):
            h = self.update(h, adj, step)
            if self.concat_hidden:
                gs.append(self.readout(h, h0, step))
            self.atoms_list.append(h)
        if self.concat_hidden:
            return F.concat(gs, axis=1)
        if self.sum_readout:
            return F.sum_(h, axis=1)          # ggnn_dev.py:167
        return self.readout(h, h0, 0)

    def get_atom_array(self, step=-1):
        return self.atoms_list[step]


# ----------------------------------------------------------------------------
# models/relgcn.py:32-73
# ----------------------------------------------------------------------------
class RelGCN(object):
    def __init__(self, p, out_channels=64, num_edge_type=4, ch_list=None,
                 n_atom_types=MAX_ATOMIC_NUM, input_type="int", scale_adj=None):
        if ch_list is None:
            ch_list = [16, 128, 64]
        if input_type not in ("int", "float"):
            raise ValueError("[ERROR] Unexpected value input type={}".format(input_type))
        self.p, self.scale_adj, self.input_type = p, scale_adj, input_type
        self.convs = [RelGCNUpdate(p.sub("rgcn_convs/%d" % i), ch_list[i], ch_list[i + 1], num_edge_type)
                      for i in range(len(ch_list) - 1)]
        self.readout = GGNNReadout(p.sub("rgcn_readout"), out_channels, ch_list[-1],
                                   nobias=True, activation="tanh")
        self.atoms = None

    def __call__(self, h, adj):
        if self.input_type == "int":
            h = F.embed_id(h, self.p["embed/W"])                                       # :67
        else:
            q = self.p.sub("embed")
            h = F.graph_linear(F.as_var(h), q["W"], q["b"])
        if self.scale_adj:
            adj = rescale_adj(adj)                                                     # :68-69
        for conv in self.convs:
            h = F.tanh(conv(h, adj))                                                   # :70-71
        self.atoms = h
        return self.readout(h)                                                         # :72

    def get_atom_array(self):
        return self.atoms


# ----------------------------------------------------------------------------
# Fine-grained co-attention.
#   models/coattention/nie_coattention.py:312-396          (NieFineCoattention)
#   models/coattention/vqa_parallel_coattention.py:13-103  (VQAParallelCoattention)
#   models/coattention/PoolingFineCoattention.py:13-83     (PoolingFineCoattention)
# ----------------------------------------------------------------------------
def _energy(p, act, query, key):
    """compute_attention (nie_coattention.py:371-396): act(Bilinear(key, query))
    over tile-materialised (mb*Nq*Nk, H) operands -> (mb, Nq, Nk)."""
    mb, nq, hd = query.shape
    nk = key.shape[1]
    q = F.reshape(F.tile(F.expand_dims(query, 2), (1, 1, nk, 1)), (mb * nq * nk, hd))
    k = F.reshape(F.tile(F.expand_dims(key, 1), (1, nq, 1, 1)), (mb * nq * nk, hd))
    e = p.sub("energy_layer")
    y = F.bilinear(k, q, e["W"], e["V1"], e["V2"], e["b"])
    return F.reshape(act(y), (mb, nq, nk))


class NieFineCoattention(object):
    default_activation = "identity"   # nie_coattention.py:316

    def __init__(self, p, hidden_dim, out_dim, head, activation=None):
        self.p, self.out_dim = p, out_dim
        self.act = ACT[activation or self.default_activation]

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        p = self.p
        C = _energy(p, self.act, query=atoms_2, key=atoms_1)                           # (mb, N2, N1) :344
        L_2 = F.softmax(C, axis=1)                                                     # :347
        L_1 = F.softmax(F.transpose(C, (0, 2, 1)), axis=1)                             # :349
        lt_1 = F.graph_linear(atoms_1, p["lt_layer_1/W"])                              # :352
        lt_2 = F.graph_linear(atoms_2, p["lt_layer_2/W"])                              # :354
        H_1 = F.tanh(F.add(lt_1, F.matmul(L_1, lt_2)))                                 # :356-358
        H_2 = F.tanh(F.add(lt_2, F.matmul(L_2, lt_1)))                                 # :361-362
        attn_1 = F.softmax(F.graph_linear(H_1, p["attention_layer_1/W"]))              # :364
        attn_2 = F.softmax(F.graph_linear(H_2, p["attention_layer_2/W"]))              # :366
        j1 = F.graph_linear(atoms_1, p["j_layer/W"], p["j_layer/b"])
        j2 = F.graph_linear(atoms_2, p["j_layer/W"], p["j_layer/b"])
        c1 = F.sum_(F.mul(F.tile(attn_1, (1, 1, self.out_dim)), j1), axis=1)           # :368
        c2 = F.sum_(F.mul(F.tile(attn_2, (1, 1, self.out_dim)), j2), axis=1)           # :369
        return c1, c2


class VQAParallelCoattention(NieFineCoattention):
    default_activation = "tanh"       # vqa_parallel_coattention.py:23


class AlternatingCoattention(object):
    """alternating_coattention.py:14-86."""

    def __init__(self, p, hidden_dim, out_dim, head, weight_tying=True):
        self.p, self.out_dim, self.weight_tying = p, out_dim, weight_tying

    def compute_attention(self, query, key, focus):
        idx = 0 if self.weight_tying else focus - 1
        q1, q2 = self.p.sub("energy_layers_1/%d" % idx), self.p.sub("energy_layers_2/%d" % idx)
        n = key.shape[1]
        query = F.tile(F.expand_dims(query, 1), (1, n, 1))                             # :73-75
        energy = F.tanh(F.graph_linear(F.concat((query, key), axis=2), q1["W"], q1["b"]))   # :77
        return F.softmax(F.graph_linear(energy, q2["W"], q2["b"]), axis=1)            # :78-79

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        j = self.p.sub("j_layer")
        attn_1 = F.tile(self.compute_attention(g_2, atoms_1, 1), (1, 1, self.out_dim))
        c1 = F.sum_(F.mul(attn_1, F.graph_linear(atoms_1, j["W"], j["b"])), axis=1)    # :44-47
        attn_2 = F.tile(self.compute_attention(c1, atoms_2, 2), (1, 1, self.out_dim))
        c2 = F.sum_(F.mul(attn_2, F.graph_linear(atoms_2, j["W"], j["b"])), axis=1)    # :50-54
        return c1, c2


class ParallelCoattention(object):
    """parallel_coattention.py:12-83."""

    def __init__(self, p, hidden_dim, out_dim, head, activation="tanh", weight_tying=True):
        self.p, self.hidden_dim, self.out_dim, self.head = p, hidden_dim, out_dim, head
        self.act, self.weight_tying = ACT[activation], weight_tying

    def compute_attention(self, query, key, focus):
        e = self.p.sub("energy_layers/%d" % (0 if self.weight_tying else focus - 1))
        mb, n, _ = key.shape
        query = F.reshape(F.tile(F.expand_dims(query, 1), (1, n, 1)), (mb * n, self.out_dim))
        key = F.reshape(key, (mb * n, self.hidden_dim))
        energy = self.act(F.bilinear(key, query, e["W"], e["V1"], e["V2"], e["b"]))    # :79
        return F.reshape(energy, (mb, n, self.head))

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        j = self.p.sub("j_layer")
        attn_1 = F.tile(self.compute_attention(g_2, atoms_1, 1), (1, 1, self.out_dim))
        c1 = F.sum_(F.mul(attn_1, F.graph_linear(atoms_1, j["W"], j["b"])), axis=1)
        attn_2 = F.tile(self.compute_attention(g_1, atoms_2, 2), (1, 1, self.out_dim))
        c2 = F.sum_(F.mul(attn_2, F.graph_linear(atoms_2, j["W"], j["b"])), axis=1)
        return c1, c2


class CircularParallelCoattention(object):
    """parallel_coattention.py:86-187."""

    def __init__(self, p, hidden_dim, out_dim, activation="tanh"):
        self.p, self.out_dim, self.act = p, out_dim, ACT[activation]

    def _corr(self, left_x, right_x):                                                  # :160-187, as hole.py:28-50
        zeros = lambda v: F.const(np.zeros_like(v.data))
        lr, li = F.fft((left_x, zeros(left_x)))
        rr, ri = F.fft((right_x, zeros(right_x)))
        out, _ = F.ifft((F.add(F.mul(lr, rr), F.mul(li, ri)), F.sub(F.mul(lr, ri), F.mul(li, rr))))
        return out

    def _side(self, atoms, query):
        j = self.p.sub("j_layer")
        atoms = F.graph_linear(atoms, j["W"], j["b"])                                  # :124
        mb, n, _ = atoms.shape
        q = F.reshape(F.tile(F.expand_dims(query, 1), (1, n, 1)), (mb * n, self.out_dim))
        k = F.reshape(atoms, (mb * n, self.out_dim))
        attn = F.reshape(self.act(self._corr(k, q)), (mb, n, self.out_dim))            # :156-157
        return F.sum_(F.mul(attn, atoms), axis=1)

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        return self._side(atoms_1, g_2), self._side(atoms_2, g_1)


class GlobalCoattention(object):
    """global_coattention.py:9-73."""

    def __init__(self, p, hidden_dim, out_dim, weight_tying=True):
        self.p, self.out_dim, self.weight_tying = p, out_dim, weight_tying

    def compute_attention(self, query, key, focus):
        q = self.p.sub("att_layers/%d" % (0 if self.weight_tying else focus - 1))
        mb, n, hd = key.shape
        query = F.reshape(F.tile(F.expand_dims(query, 1), (1, n, 1)), (mb * n, hd))    # :62-64
        key = F.reshape(key, (mb * n, hd))
        energy = F.sigmoid(F.linear(F.concat((key, query), axis=-1), q["W"], q["b"]))  # :67
        return F.reshape(energy, (mb, n, self.out_dim))

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        lt = self.p.sub("lt_layer")
        g1, g2 = F.mean(atoms_1, axis=1), F.mean(atoms_2, axis=1)                      # :36-37
        c1 = F.sum_(F.mul(self.compute_attention(g2, atoms_1, 1), F.graph_linear(atoms_1, lt["W"], lt["b"])), axis=1)
        c2 = F.sum_(F.mul(self.compute_attention(g1, atoms_2, 2), F.graph_linear(atoms_2, lt["W"], lt["b"])), axis=1)
        return c1, c2


class NeuralCoattention(object):
    """neural_coattention.py:8-71."""

    def __init__(self, p, hidden_dim, out_dim, activation="relu", weight_tying=True):
        self.p, self.out_dim, self.act, self.weight_tying = p, out_dim, ACT[activation], weight_tying

    def compute_attention(self, query, key, focus):
        q = self.p.sub("att_layers/%d" % (0 if self.weight_tying else focus - 1))
        query = F.expand_dims(query, 1)                                                # :64
        context = self.act(F.graph_linear(query, q["W"], q["b"]))                      # :66
        doc = self.act(F.graph_linear(key, q["W"], q["b"]))                            # :67
        energy = F.sigmoid(F.matmul(doc, F.transpose(context, (0, 2, 1))))             # :68
        return energy, doc

    def _side(self, query, key, focus):
        attn, doc = self.compute_attention(query, key, focus)
        return F.sum_(F.mul(F.tile(attn, (1, 1, self.out_dim)), doc), axis=1)

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        c1 = self._side(F.mean(atoms_2, axis=1), atoms_1, 1)
        c2 = self._side(F.mean(atoms_1, axis=1), atoms_2, 2)
        return c1, c2


def vector_coattn_shapes(kind, hidden_dim, out_dim, head=1):
    s = {"j_layer/W": (out_dim, hidden_dim), "j_layer/b": (out_dim,)}
    if kind == "alter":
        s.update({"energy_layers_1/0/W": (head, hidden_dim + out_dim), "energy_layers_1/0/b": (head,),
                  "energy_layers_2/0/W": (1, head), "energy_layers_2/0/b": (1,)})
    elif kind == "para":
        s.update({"energy_layers/0/W": (hidden_dim, out_dim, head), "energy_layers/0/V1": (hidden_dim, head),
                  "energy_layers/0/V2": (out_dim, head), "energy_layers/0/b": (head,)})
    elif kind == "global":
        s = {"lt_layer/W": (out_dim, hidden_dim), "lt_layer/b": (out_dim,),
             "att_layers/0/W": (out_dim, 2 * hidden_dim), "att_layers/0/b": (out_dim,)}
    elif kind == "neural":
        s = {"att_layers/0/W": (out_dim, hidden_dim), "att_layers/0/b": (out_dim,)}
    return s


class FourierFineCoattention(NieFineCoattention):
    """nie_coattention.py:399-515: the energy map is act(Bilinear(Re fft(key), Re fft(query)) + Bilinear(Im fft(key), Im fft(query)))
    with the FFT over the hidden axis (:484-491,:505-515); the head path sees the atoms themselves."""

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        p = self.p
        C = self._energy(query=atoms_2, key=atoms_1)
        L_2 = F.softmax(C, axis=1)
        L_1 = F.softmax(F.transpose(C, (0, 2, 1)), axis=1)
        lt_1 = F.graph_linear(atoms_1, p["lt_layer_1/W"])
        lt_2 = F.graph_linear(atoms_2, p["lt_layer_2/W"])
        H_1 = F.tanh(F.add(lt_1, F.matmul(L_1, lt_2)))
        H_2 = F.tanh(F.add(lt_2, F.matmul(L_2, lt_1)))
        attn_1 = F.softmax(F.graph_linear(H_1, p["attention_layer_1/W"]))
        attn_2 = F.softmax(F.graph_linear(H_2, p["attention_layer_2/W"]))
        j1 = F.graph_linear(atoms_1, p["j_layer/W"], p["j_layer/b"])
        j2 = F.graph_linear(atoms_2, p["j_layer/W"], p["j_layer/b"])
        c1 = F.sum_(F.mul(F.tile(attn_1, (1, 1, self.out_dim)), j1), axis=1)
        c2 = F.sum_(F.mul(F.tile(attn_2, (1, 1, self.out_dim)), j2), axis=1)
        return c1, c2

    def _energy(self, query, key):
        mb, nq, hd = query.shape
        nk = key.shape[1]
        zeros = lambda v: F.const(np.zeros_like(v.data))
        q_re, q_im = F.fft((query, zeros(query)))                                      # :505-515
        k_re, k_im = F.fft((key, zeros(key)))
        tq = lambda v: F.reshape(F.tile(F.expand_dims(v, 2), (1, 1, nk, 1)), (mb * nq * nk, hd))
        tk = lambda v: F.reshape(F.tile(F.expand_dims(v, 1), (1, nq, 1, 1)), (mb * nq * nk, hd))
        e = self.p.sub("energy_layer")
        bil = lambda a, b: F.bilinear(a, b, e["W"], e["V1"], e["V2"], e["b"])
        y = F.add(bil(tk(k_re), tq(q_re)), bil(tk(k_im), tq(q_im)))                    # :497
        return F.reshape(self.act(y), (mb, nq, nk))


class DeepNieFineCoattention(NieFineCoattention):
    """nie_coattention.py:13-104 (one GraphLinear(H,H) before the head projection), :107-203 (VeryDeep, two),
    :206-309 (ExtremeDeep, three).  The energy map sees the ORIGINAL atoms (:42,:144); the head projections and j_layer see
    the transformed ones (:50-55,:73-74)."""
    n_lt_layers = 1

    def _prev(self, side, atoms):
        for i in range(self.n_lt_layers):
            name = "prev_lt_layer_%d" % side if self.n_lt_layers == 1 else "prev_lt_layers_%d/%d" % (side, i)
            q = self.p.sub(name)
            atoms = F.graph_linear(atoms, q["W"], q["b"])
        return atoms

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        p = self.p
        C = _energy(p, self.act, query=atoms_2, key=atoms_1)
        L_2 = F.softmax(C, axis=1)
        L_1 = F.softmax(F.transpose(C, (0, 2, 1)), axis=1)
        atoms_1 = self._prev(1, atoms_1)
        lt_1 = F.graph_linear(atoms_1, p["lt_layer_1/W"])
        atoms_2 = self._prev(2, atoms_2)
        lt_2 = F.graph_linear(atoms_2, p["lt_layer_2/W"])
        H_1 = F.tanh(F.add(lt_1, F.matmul(L_1, lt_2)))
        H_2 = F.tanh(F.add(lt_2, F.matmul(L_2, lt_1)))
        attn_1 = F.softmax(F.graph_linear(H_1, p["attention_layer_1/W"]))
        attn_2 = F.softmax(F.graph_linear(H_2, p["attention_layer_2/W"]))
        j1 = F.graph_linear(atoms_1, p["j_layer/W"], p["j_layer/b"])
        j2 = F.graph_linear(atoms_2, p["j_layer/W"], p["j_layer/b"])
        c1 = F.sum_(F.mul(F.tile(attn_1, (1, 1, self.out_dim)), j1), axis=1)
        c2 = F.sum_(F.mul(F.tile(attn_2, (1, 1, self.out_dim)), j2), axis=1)
        return c1, c2


class VeryDeepNieFineCoattention(DeepNieFineCoattention):
    n_lt_layers = 2


class ExtremeDeepNieFineCoattention(DeepNieFineCoattention):
    n_lt_layers = 3


class PoolingFineCoattention(object):
    def __init__(self, p, hidden_dim, out_dim, activation="tanh"):
        self.p, self.out_dim, self.act = p, out_dim, ACT[activation]

    def __call__(self, atoms_1, g_1, atoms_2, g_2):
        p = self.p
        energy = _energy(p, self.act, query=atoms_2, key=atoms_1)                      # (mb, N2, N1) :40
        attn_1 = F.softmax(F.mean(energy, axis=1), axis=1)                             # :42-44
        attn_2 = F.softmax(F.mean(energy, axis=2), axis=1)                             # :48-50
        attn_1 = F.tile(F.expand_dims(attn_1, 2), (1, 1, self.out_dim))
        attn_2 = F.tile(F.expand_dims(attn_2, 2), (1, 1, self.out_dim))
        j1 = F.graph_linear(atoms_1, p["j_layer/W"], p["j_layer/b"])
        j2 = F.graph_linear(atoms_2, p["j_layer/W"], p["j_layer/b"])
        return F.sum_(F.mul(attn_1, j1), axis=1), F.sum_(F.mul(attn_2, j2), axis=1)    # :54-56


# ----------------------------------------------------------------------------
# models/link_prediction/hole.py:53-91 (= models/mlp.py:113-151)
# ----------------------------------------------------------------------------
class HolE(object):
    def __init__(self, p, out_dim, hidden_dims=(32, 16), activation="relu",
                 layers_name="hidden_layers"):
        self.p, self.n_hidden, self.act, self.layers_name = p, len(hidden_dims), ACT[activation], layers_name

    def circular_correlation(self, left_x, right_x):
        zeros = lambda v: F.const(np.zeros_like(v.data))
        lr, li = F.fft((left_x, zeros(left_x)))                                        # :38-40
        rr, ri = F.fft((right_x, zeros(right_x)))                                      # :42-44
        pr = F.add(F.mul(lr, rr), F.mul(li, ri))                                       # :46
        pi = F.sub(F.mul(lr, ri), F.mul(li, rr))                                       # :47
        out, _ = F.ifft((pr, pi))                                                      # :49
        return out

    def __call__(self, left_x, right_x):
        h = self.circular_correlation(left_x, right_x)
        for i in range(self.n_hidden):
            q = self.p.sub("%s/%d" % (self.layers_name, i))
            h = self.act(F.linear(h, q["W"], q["b"]))
        q = self.p.sub("l_out")
        return F.linear(h, q["W"], q["b"])


# ----------------------------------------------------------------------------
# models/gin.py:58-190 (GINUpdate, GIN; dropout_ratio = 0)
# ----------------------------------------------------------------------------
class GINUpdate(object):
    def __init__(self, p, hidden_dim=16):
        self.p = p

    def __call__(self, h, adj):
        a = F.sum_(F.as_var(adj), axis=1)                                              # :88
        sum_h = F.add(F.matmul(a, h), h)                                               # :91-94
        q1, q2 = self.p.sub("linear_g1"), self.p.sub("linear_g2")
        new_h = F.relu(F.graph_linear(sum_h, q1["W"], q1["b"]))                        # :97
        return F.relu(F.graph_linear(new_h, q2["W"], q2["b"]))                         # :103


class GIN(object):
    def __init__(self, p, out_dim, hidden_dim=16, n_layers=4, concat_hidden=False, weight_tying=True, activation="identity"):
        self.p, self.concat_hidden, self.weight_tying = p, concat_hidden, weight_tying
        self.n_message_layers = 1 if weight_tying else n_layers                        # :131,:154 (the loop bound)
        self.readouts = [GGNNReadout(p.sub("readout_layers/%d" % i), out_dim, hidden_dim, activation=activation, activation_agg=activation)
                         for i in range(n_layers if concat_hidden else 1)]
        self.updates = [GINUpdate(p.sub("update_layers/%d" % i), hidden_dim) for i in range(self.n_message_layers)]
        self.atoms = None

    def __call__(self, atom_array, adj, is_real_node=None):
        a = np.asarray(getattr(atom_array, "data", atom_array))
        h = F.embed_id(a, self.p["embed/W"]) if a.ndim <= 2 else F.as_var(atom_array)  # :148-151
        h0 = F.copy(h)
        gs = []
        for step in range(self.n_message_layers):
            h = self.updates[0 if self.weight_tying else step](h, adj)
            if self.concat_hidden:
                gs.append(self.readouts[step](h, h0, is_real_node))
        self.atoms = h
        return F.concat(gs, axis=1) if self.concat_hidden else self.readouts[0](h, h0, is_real_node)


def gin_shapes(out_dim, hidden_dim, n_layers, concat_hidden=False, weight_tying=True, n_atom_types=MAX_ATOMIC_NUM):
    s = {"embed/W": (n_atom_types, hidden_dim)}
    for i in range(1 if weight_tying else n_layers):
        for l in ("linear_g1", "linear_g2"):
            s["update_layers/%d/%s/W" % (i, l)], s["update_layers/%d/%s/b" % (i, l)] = (hidden_dim, hidden_dim), (hidden_dim,)
    for i in range(n_layers if concat_hidden else 1):
        for l in ("i_layer", "j_layer"):
            s["readout_layers/%d/%s/W" % (i, l)], s["readout_layers/%d/%s/b" % (i, l)] = (out_dim, 2 * hidden_dim), (out_dim,)
    return s


# ----------------------------------------------------------------------------
# models/coattention/bimpm.py:17-197 -- bilateral multi-perspective matching (--attn bimpm, train_binary.py:253-256)
# ----------------------------------------------------------------------------
class BiMPM(object):
    """Restated op by op, including what the file actually computes rather than what its comments say: `mp_matching_func`
    multiplies (head x hidden) by (hidden x head) and keeps COLUMN 0 (:79-81), i.e. perspective k of v1 is matched against
    perspective 0 of v2; `out_dim` is unused (the output has num_match * head columns)."""

    def __init__(self, p, hidden_dim, out_dim, head, with_max_pool=True, with_att_mean=True, with_att_max=True):
        self.p, self.hidden_dim, self.head = p, hidden_dim, head
        self.flags = (with_max_pool, with_att_mean, with_att_max)

    def _match(self, v1, v2, w):                                                       # :50-82
        mb, n1, _ = v1.shape
        wt = F.expand_dims(F.expand_dims(F.transpose(w, (1, 0)), 0), 0)               # (1,1,hidden,head)
        tw = F.tile(wt, (mb, n1, 1, 1))
        a = F.mul(tw, F.stack([v1] * self.head, axis=3))
        b = F.mul(tw, F.stack([v2] * self.head, axis=3))
        sim = F.matmul(F.transpose(F.normalize(a, axis=2), (0, 1, 3, 2)), F.normalize(b, axis=2))
        return F.getitem(sim, (slice(None), slice(None), slice(None), 0))

    def _match_pairwise(self, v1, v2, w):                                              # :84-108
        mb, n1, _ = v1.shape
        n2 = v2.shape[1]
        we = F.expand_dims(F.expand_dims(w, 0), 2)                                     # (1,head,1,hidden)
        a = F.mul(F.tile(we, (mb, 1, n1, 1)), F.stack([v1] * self.head, axis=1))
        b = F.mul(F.tile(we, (mb, 1, n2, 1)), F.stack([v2] * self.head, axis=1))
        sim = F.matmul(F.normalize(a, axis=3), F.transpose(F.normalize(b, axis=3), (0, 1, 3, 2)))
        return F.transpose(sim, (0, 2, 3, 1))                                          # (mb,N1,N2,head)

    @staticmethod
    def _div(n, d, eps=1e-4):                                                          # :125-127
        dd = F.maximum(d, np.ones_like(d.data) * eps)
        return F.div(n, F.Var(np.broadcast_to(dd.data, n.shape).copy(), (dd,), lambda g: (F._unbroadcast(g, dd.shape),)))

    def __call__(self, atoms_1, g1, atoms_2, g2):
        atoms_1, atoms_2 = F.as_var(atoms_1), F.as_var(atoms_2)
        mb, n1, _ = atoms_1.shape
        n2 = atoms_2.shape[1]
        with_max_pool, with_att_mean, with_att_max = self.flags
        mv1, mv2 = [], []
        if with_max_pool:                                                              # :136-146
            mv = self._match_pairwise(atoms_1, atoms_2, self.p["max_pooling_W"])
            mv1.append(F.max_(mv, axis=2))
            mv2.append(F.max_(mv, axis=1))
        if with_att_mean or with_att_max:                                              # :148-160
            att = F.matmul(F.normalize(atoms_1, axis=2), F.transpose(F.normalize(atoms_2, axis=2), (0, 2, 1)))
            att_e = F.tile(F.expand_dims(att, 3), (1, 1, 1, self.hidden_dim))
            att_atoms2 = F.mul(F.tile(F.expand_dims(atoms_2, 1), (1, n1, 1, 1)), att_e)
            att_atoms1 = F.mul(F.tile(F.expand_dims(atoms_1, 2), (1, 1, n2, 1)), att_e)
            if with_att_mean:                                                          # :162-172
                m2 = self._div(F.sum_(att_atoms2, axis=2), F.expand_dims(F.sum_(att, axis=2), 2))
                m1 = self._div(F.sum_(att_atoms1, axis=1), F.transpose(F.expand_dims(F.sum_(att, axis=1), 1), (0, 2, 1)))
                mv1.append(self._match(atoms_1, m2, self.p["att_mean_W"]))
                mv2.append(self._match(atoms_2, m1, self.p["att_mean_W"]))
            if with_att_max:                                                           # :174-186
                mv1.append(self._match(atoms_1, F.max_(att_atoms2, axis=2), self.p["att_max_W"]))
                mv2.append(self._match(atoms_2, F.max_(att_atoms1, axis=1), self.p["att_max_W"]))
        return F.sum_(F.concat(mv1, axis=2), axis=1), F.sum_(F.concat(mv2, axis=2), axis=1)     # :188-197 (aggr = F.sum)


def bimpm_shapes(hidden_dim, head):
    return {"max_pooling_W": (head, hidden_dim), "att_mean_W": (head, hidden_dim), "att_max_W": (head, hidden_dim)}


# ----------------------------------------------------------------------------
# models/models/nfp.py:15-181 -- Neural Fingerprint encoder (the default --method of train_binary.py:319)
# ----------------------------------------------------------------------------
class NFPUpdate(object):
    def __init__(self, p, in_channels, out_channels, max_degree=6):
        self.p, self.n_deg = p, max_degree + 1

    def __call__(self, h, adj, deg_conds):
        fv = F.matmul(F.as_var(adj), h)                                                # :44 neighbour sum (adj holds the self loop)
        zero = np.zeros(fv.shape, dtype=fv.dtype)                                      # :48-51
        out_h = None
        for d, cond in enumerate(deg_conds):                                           # :53-57 one GraphLinear per degree, masked input
            q = self.p.sub("graph_linears/%d" % d)
            y = F.graph_linear(F.where(cond, fv, zero), q["W"], q["b"])
            out_h = y if out_h is None else F.add(out_h, y)
        return F.sigmoid(out_h)                                                        # :60


class NFPReadout(object):
    def __init__(self, p, in_channels, out_size):
        self.p = p

    def __call__(self, h):
        q = self.p.sub("output_weight")
        i = F.softmax(F.graph_linear(h, q["W"], q["b"]), axis=2)                       # :88-89 softmax along the channel axis
        return F.sum_(i, axis=1)                                                       # :90 sum along the atom axis


class NFP(object):
    def __init__(self, p, out_dim, hidden_dim=16, n_layers=4, max_degree=6):
        self.p, self.num_degree_type = p, max_degree + 1
        self.layers = [NFPUpdate(p.sub("layers/%d" % i), hidden_dim, hidden_dim, max_degree) for i in range(n_layers)]
        self.read_out_layers = [NFPReadout(p.sub("read_out_layers/%d" % i), hidden_dim, out_dim) for i in range(n_layers)]
        self.atoms = None

    def __call__(self, atom_array, adj):
        a = np.asarray(getattr(atom_array, "data", atom_array))
        h = F.embed_id(a, self.p["embed/W"]) if a.dtype.kind == "i" else F.as_var(atom_array)      # :142-146
        adj_array = np.asarray(getattr(adj, "data", adj))
        degree_mat = adj_array.sum(axis=1)                                             # :152
        deg_conds = [np.broadcast_to(((degree_mat - degree) == 0)[:, :, None], h.shape)
                     for degree in range(1, self.num_degree_type + 1)]                  # :154-156
        g = None
        for update, readout in zip(self.layers, self.read_out_layers):                 # :158-163
            h = update(h, adj, deg_conds)
            dg = readout(h)
            g = dg if g is None else F.add(g, dg)
        self.atoms = h
        return g

    def get_atom_array(self):
        return self.atoms


def nfp_shapes(out_dim, hidden_dim, n_layers, max_degree=6, n_atom_types=MAX_ATOMIC_NUM):
    s = {"embed/W": (n_atom_types, hidden_dim)}
    for i in range(n_layers):
        for d in range(max_degree + 1):
            s["layers/%d/graph_linears/%d/W" % (i, d)], s["layers/%d/graph_linears/%d/b" % (i, d)] = (hidden_dim, hidden_dim), (hidden_dim,)
        s["read_out_layers/%d/output_weight/W" % i], s["read_out_layers/%d/output_weight/b" % i] = (out_dim, hidden_dim), (out_dim,)
    return s


# ----------------------------------------------------------------------------
# models/mlp.py:20-110,154-197 -- the other link-prediction heads
# ----------------------------------------------------------------------------
def _stack(p, names, n_hidden, act, h):
    for i in range(n_hidden):
        q = p.sub("%s/%d" % (names, i))
        h = act(F.linear(h, q["W"], q["b"]))
    q = p.sub("l_out")
    return F.linear(h, q["W"], q["b"])


class MLP(object):
    """models/mlp.py:20-45."""

    def __init__(self, p, out_dim, hidden_dims=(32, 16), activation="relu"):
        self.p, self.n_hidden, self.act = p, len(hidden_dims), ACT[activation]

    def __call__(self, x):
        return _stack(self.p, "layers", self.n_hidden, self.act, x)                    # :40-44


class SymMLP(MLP):
    """models/mlp.py:95-110."""

    def __call__(self, left_x, right_x):
        h = F.concat((F.add(left_x, right_x), F.mul(left_x, right_x)), axis=1)        # :105
        return _stack(self.p, "layers", self.n_hidden, self.act, h)


class NTN(object):
    """models/mlp.py:47-74: links.Bilinear (with V1, V2, b) -> Linear stack; no activation after the bilinear layer."""

    def __init__(self, p, left_dim, right_dim, out_dim, ntn_out_dim=8, hidden_dims=(16,), activation="relu"):
        self.p, self.n_hidden, self.act = p, len(hidden_dims), ACT[activation]

    def __call__(self, left_x, right_x):
        q = self.p.sub("ntn_layer")
        h = F.bilinear(left_x, right_x, q["W"], q["V1"], q["V2"], q["b"])              # :67
        return _stack(self.p, "mlp_layers", self.n_hidden, self.act, h)               # :69-72


class DistMult(object):
    """models/mlp.py:77-93 with BilinearDiag (:154-197): the (L, R, out) tensor of diagonal slices is built from
    `self.W.data` (:186-192), i.e. as a constant -- the diagonal weights receive no gradient in the reference."""

    def __init__(self, p, left_dim, right_dim, out_dim, dm_out_dim=8, hidden_dims=(16,), activation="relu"):
        self.p, self.n_hidden, self.act = p, len(hidden_dims), ACT[activation]

    def __call__(self, left_x, right_x):
        Wd = self.p.sub("dm_layer")["W"].data                                          # (out, L)
        W_mat = np.array([np.diag(v) for v in Wd], dtype=np.float32).transpose(1, 2, 0)  # :188-190 -> (L, R, out), float32 as there
        h = F.bilinear(left_x, right_x, F.const(W_mat))                                # :176 bilinear.bilinear(e1, e2, W_mat)
        return _stack(self.p, "mlp_layers", self.n_hidden, self.act, h)


def head_shapes(kind, in_dim, out_dim, hidden_dims, mid=8):
    """Chainer parameter names/shapes of the heads above (kind in mlp / symmlp / ntn / distmult)."""
    names = "layers" if kind in ("mlp", "symmlp") else "mlp_layers"
    d = {"mlp": 2 * in_dim, "symmlp": 2 * in_dim, "ntn": mid, "distmult": mid}[kind]
    s = {}
    if kind == "ntn":
        s.update({"ntn_layer/W": (in_dim, in_dim, mid), "ntn_layer/V1": (in_dim, mid), "ntn_layer/V2": (in_dim, mid), "ntn_layer/b": (mid,)})
    if kind == "distmult":
        s["dm_layer/W"] = (mid, in_dim)
    for i, hdim in enumerate(hidden_dims):
        s["%s/%d/W" % (names, i)], s["%s/%d/b" % (names, i)] = (hdim, d), (hdim,)
        d = hdim
    s["l_out/W"], s["l_out/b"] = (out_dim, d), (out_dim,)
    return s


def apply_hooks(grads, params, max_norm=0.0, l2_rate=0.0, l1_rate=0.0):
    """chainer.optimizer hooks in the order train_binary.py:537-543 adds them, on dicts of arrays (in place on copies):
    GradientClipping (rate = threshold / sqrt(sum of squared norms over ALL parameters); scale when rate < 1),
    WeightDecay (g += rate * p), Lasso (g += rate * sign(p))."""
    g = {k: np.array(v, dtype=np.float64) for k, v in grads.items()}
    if max_norm > 0:
        norm = np.sqrt(sum(float((v * v).sum()) for v in g.values()))
        rate = max_norm / norm
        if rate < 1:
            for k in g:
                g[k] *= rate
    if l2_rate > 0:
        for k in g:
            g[k] += l2_rate * params[k]
    if l1_rate > 0:
        for k in g:
            g[k] += l1_rate * np.sign(params[k])
    return g


# ----------------------------------------------------------------------------
# train_binary.py:84-118 (pair composition) and :524 (loss)
# ----------------------------------------------------------------------------
class GraphConvPredictorForPair(object):
    """train_binary.py:84-118; `first_last_atoms=True`: train_ddi_modify_eval3.py:110-134 (default flags), where the co-attention
    sees [atoms after the first step || atoms after the last step], 2*hidden wide."""

    def __init__(self, graph_conv, attn=None, mlp=None, first_last_atoms=False):
        self.graph_conv, self.attn, self.mlp, self.first_last_atoms = graph_conv, attn, mlp, first_last_atoms

    def _atoms(self):
        if not self.first_last_atoms:
            return self.graph_conv.get_atom_array()
        return F.concat([self.graph_conv.get_atom_array(0), self.graph_conv.get_atom_array(-1)], axis=2)   # eval3 :117-120

    def __call__(self, atoms_1, adjs_1, atoms_2, adjs_2):
        g1 = self.graph_conv(atoms_1, adjs_1)
        a1 = self._atoms()
        g2 = self.graph_conv(atoms_2, adjs_2)
        a2 = self._atoms()
        if self.attn is not None:
            g1, g2 = self.attn(a1, g1, a2, g2)
        if type(self.mlp) is MLP:                                                      # train_binary.py:98-100
            return self.mlp(F.concat((g1, g2), axis=-1))
        return self.mlp(g1, g2)


def loss_and_grads(predictor, table, inputs, labels):
    """One Classifier iteration: sigmoid-CE loss + backward (train_binary.py:524-530)."""
    for v in table.values():
        v.grad = None
    logits = predictor(*inputs)
    loss = F.sigmoid_cross_entropy(logits, labels)
    loss.backward()
    return loss.data, logits.data, grads_of(table)


# ----------------------------------------------------------------------------
# parameter shapes (Chainer names) + Chainer-like initialisation
# ----------------------------------------------------------------------------
def _gru_shapes(pre, H):
    s = {}
    for n in ("W_r", "W_z", "W"):
        s[pre + n + "/W"], s[pre + n + "/b"] = (H, 2 * H), (H,)
    for n in ("U_r", "U_z", "U"):
        s[pre + n + "/W"], s[pre + n + "/b"] = (H, H), (H,)
    return s


def ggnn_shapes(out_dim, hidden_dim, n_layers, concat_hidden=False, weight_tying=True, E=4,
                n_atom_types=MAX_ATOMIC_NUM):
    H = hidden_dim
    s = {"embed/W": (n_atom_types, H)}
    for i in range(1 if weight_tying else n_layers):
        pre = "update_layers/%d/" % i
        s[pre + "graph_linear/W"], s[pre + "graph_linear/b"] = (E * H, H), (E * H,)
        s.update(_gru_shapes(pre + "update_layer/", H))
    for i in range(n_layers if concat_hidden else 1):
        pre = "readout_layers/%d/" % i
        for n in ("i_layer", "j_layer"):
            s[pre + n + "/W"], s[pre + n + "/b"] = (out_dim, 2 * H), (out_dim,)
    return s


def ggnn_mono_shapes(out_dim, hidden_dim, n_layers, concat_hidden=False, weight_tying=True, E=4,
                     n_atom_types=MAX_ATOMIC_NUM):
    H = hidden_dim
    s = {"embed/W": (n_atom_types, H)}
    for i in range(1 if weight_tying else n_layers):
        s["message_layers/%d/W" % i], s["message_layers/%d/b" % i] = (E * H, H), (E * H,)
    s.update(_gru_shapes("update_layer/", H))
    for i in range(n_layers if concat_hidden else 1):
        s["i_layers/%d/W" % i], s["i_layers/%d/b" % i] = (out_dim, 2 * H), (out_dim,)
        s["j_layers/%d/W" % i], s["j_layers/%d/b" % i] = (out_dim, H), (out_dim,)
    return s


def relgcn_shapes(out_channels, ch_list, E=4, n_atom_types=MAX_ATOMIC_NUM):
    s = {"embed/W": (n_atom_types, ch_list[0])}
    for i in range(len(ch_list) - 1):
        pre = "rgcn_convs/%d/" % i
        s[pre + "graph_linear_self/W"], s[pre + "graph_linear_self/b"] = (ch_list[i + 1], ch_list[i]), (ch_list[i + 1],)
        s[pre + "graph_linear_edge/W"] = (ch_list[i + 1] * E, ch_list[i])
        s[pre + "graph_linear_edge/b"] = (ch_list[i + 1] * E,)
    s["rgcn_readout/i_layer/W"] = (out_channels, ch_list[-1])
    s["rgcn_readout/j_layer/W"] = (out_channels, ch_list[-1])
    return s


def coattn_shapes(hidden_dim, out_dim, head=None):
    s = {"energy_layer/W": (hidden_dim, hidden_dim, 1), "energy_layer/V1": (hidden_dim, 1),
         "energy_layer/V2": (hidden_dim, 1), "energy_layer/b": (1,),
         "j_layer/W": (out_dim, hidden_dim), "j_layer/b": (out_dim,)}
    if head is not None:
        s.update({"attention_layer_1/W": (1, head), "attention_layer_2/W": (1, head),
                  "lt_layer_1/W": (head, hidden_dim), "lt_layer_2/W": (head, hidden_dim)})
    return s


def deep_coattn_shapes(hidden_dim, out_dim, head, n_lt_layers=1):
    s = coattn_shapes(hidden_dim, out_dim, head)
    for side in (1, 2):
        for i in range(n_lt_layers):
            name = "prev_lt_layer_%d" % side if n_lt_layers == 1 else "prev_lt_layers_%d/%d" % (side, i)
            s[name + "/W"], s[name + "/b"] = (hidden_dim, hidden_dim), (hidden_dim,)
    return s


def hole_shapes(in_dim, out_dim, hidden_dims=(32, 16), layers_name="hidden_layers"):
    s, d = {}, in_dim
    for i, hdim in enumerate(hidden_dims):
        s["%s/%d/W" % (layers_name, i)], s["%s/%d/b" % (layers_name, i)] = (hdim, d), (hdim,)
        d = hdim
    s["l_out/W"], s["l_out/b"] = (out_dim, d), (out_dim,)
    return s


def init_params(shapes, rng, dtype=np.float32, bias_scale=0.1, prefix=""):
    """Chainer-like init: LeCunNormal (std 1/sqrt(fan_in)) for matrices, N(0,1)
    embedding; biases are small random instead of zero so bias paths are tested."""
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("embed/W") and len(shp) == 2 and shp[0] == MAX_ATOMIC_NUM:
            a = rng.standard_normal(shp) * 0.5
        elif len(shp) == 1:
            a = rng.standard_normal(shp) * bias_scale
        elif len(shp) == 3:
            a = rng.standard_normal(shp) / np.sqrt(shp[0])
        elif name.endswith("/V1") or name.endswith("/V2"):
            a = rng.standard_normal(shp) / np.sqrt(shp[0])
        else:
            a = rng.standard_normal(shp) / np.sqrt(shp[1])
        out[prefix + name] = a.astype(dtype)
    return out
