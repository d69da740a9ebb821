"""Independent second implementation of the hot path (Torch CPU, autograd) used ONLY
to pin the NumPy oracle: written in closed einsum form from SURVEY.md Appendix B /
section 8a, not op-by-op like the oracle, so the two can only agree if both restate
the same mathematics.  Parameters arrive as the same {chainer path: array} tables."""
import torch


def T(table, dtype=torch.float64):
    return {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in table.items()}


def lin(P, pre, x):
    return x @ P[pre + "/W"].T + P[pre + "/b"]


def message(W, b, h, adj, E=4):
    # W (E*H, H) rows c*E+e ; M[b,e,n,c] = h W_e^T + b_e ; m = sum_e A_e M_e
    Hd = h.shape[-1]
    We = W.view(Hd, E, Hd)           # [c, e, k]
    be = b.view(Hd, E)               # [c, e]
    M = torch.einsum("bnk,cek->benc", h, We) + be.T[None, :, None, :]
    return torch.einsum("beij,bejc->bic", adj, M)


def gru(P, pre, x, s):
    if s is None:
        z = torch.sigmoid(lin(P, pre + "/W_z", x))
        hb = torch.tanh(lin(P, pre + "/W", x))
        return z * hb
    r = torch.sigmoid(lin(P, pre + "/W_r", x) + lin(P, pre + "/U_r", s))
    z = torch.sigmoid(lin(P, pre + "/W_z", x) + lin(P, pre + "/U_z", s))
    hb = torch.tanh(lin(P, pre + "/W", x) + lin(P, pre + "/U", r * s))
    return z * hb + (1 - z) * s


ACT = {"identity": lambda x: x, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}


def readout_r1(P, pre, h, h0, act="identity", act_agg="identity", mask=None, nobias=False):
    h1 = torch.cat((h, h0), dim=2) if h0 is not None else h
    u = h1 @ P[pre + "/i_layer/W"].T
    v = h1 @ P[pre + "/j_layer/W"].T
    if not nobias:
        u, v = u + P[pre + "/i_layer/b"], v + P[pre + "/j_layer/b"]
    g = torch.sigmoid(u) * ACT[act](v)
    if mask is not None:
        g = g * mask[:, :, None]
    return ACT[act_agg](g.sum(dim=1))


def ggnn(P, atoms, adj, n_layers, weight_tying=True, concat_hidden=False, activation="identity", mask=None):
    """models/models/ggnn.py semantics: untied => every step is a fresh (stateless) GRU."""
    h = P["embed/W"][torch.as_tensor(atoms, dtype=torch.long)]
    h0 = h
    gs, s = [], None
    for t in range(n_layers):
        pre = "update_layers/%d" % (0 if weight_tying else t)
        m = message(P[pre + "/graph_linear/W"], P[pre + "/graph_linear/b"], h, adj)
        x = torch.cat((h, m), dim=2)
        h = gru(P, pre + "/update_layer", x, s if weight_tying else None)
        s = h
        if concat_hidden:
            gs.append(readout_r1(P, "readout_layers/%d" % t, h, h0, activation, activation, mask))
    if concat_hidden:
        return torch.cat(gs, dim=1), h
    return readout_r1(P, "readout_layers/0", h, h0, activation, activation, mask), h


def ggnn_mono(P, atoms, adj, n_layers, weight_tying=True, sum_readout=False, concat_hidden=False):
    """models/ggnn_att.py / ggnn_dev.py default path: shared stateful GRU, readout R2."""
    h = P["embed/W"][torch.as_tensor(atoms, dtype=torch.long)]
    h0 = h
    s, gs = None, []

    def ro(h, idx):
        gate = torch.sigmoid(torch.cat((h, h0), 2) @ P["i_layers/%d/W" % idx].T + P["i_layers/%d/b" % idx])
        return (gate * (h @ P["j_layers/%d/W" % idx].T + P["j_layers/%d/b" % idx])).sum(dim=1)
    for t in range(n_layers):
        pre = "message_layers/%d" % (0 if weight_tying else t)
        m = message(P[pre + "/W"], P[pre + "/b"], h, adj)
        h = gru(P, "update_layer", torch.cat((h, m), dim=2), s)
        s = h
        if concat_hidden:
            gs.append(ro(h, t))
    if concat_hidden:
        return torch.cat(gs, dim=1), h
    if sum_readout:
        return h.sum(dim=1), h
    return ro(h, 0), h


def relgcn(P, atoms, adj, ch_list, scale_adj=False):
    h = P["embed/W"][torch.as_tensor(atoms, dtype=torch.long)]
    if scale_adj:
        deg = adj.sum(dim=(1, 2))
        adj = adj / torch.where(deg != 0, deg, torch.ones_like(deg))[:, None, None, :]
    for l in range(len(ch_list) - 1):
        pre = "rgcn_convs/%d" % l
        m = message_rect(P[pre + "/graph_linear_edge/W"], P[pre + "/graph_linear_edge/b"], h, adj, ch_list[l + 1])
        h = torch.tanh(lin(P, pre + "/graph_linear_self", h) + m)
    g = readout_r1(P, "rgcn_readout", h, None, act="tanh", nobias=True)
    return g, h


def message_rect(W, b, h, adj, cout, E=4):
    We = W.view(cout, E, h.shape[-1])
    be = b.view(cout, E)
    M = torch.einsum("bnk,cek->benc", h, We) + be.T[None, :, None, :]
    return torch.einsum("beij,bejc->bic", adj, M)


def coattention(P, a1, a2, variant="fine", act="tanh"):
    W = P["energy_layer/W"][:, :, 0]
    C = torch.einsum("bjh,hk,bik->bij", a1, W, a2) \
        + (a1 @ P["energy_layer/V1"])[:, None, :, 0] + (a2 @ P["energy_layer/V2"])[:, :, 0:1] + P["energy_layer/b"]
    C = ACT[act](C)                                   # (b, N2 [i], N1 [j])
    j1 = a1 @ P["j_layer/W"].T + P["j_layer/b"]
    j2 = a2 @ P["j_layer/W"].T + P["j_layer/b"]
    if variant == "pool":
        at1 = torch.softmax(C.mean(dim=1), dim=1)
        at2 = torch.softmax(C.mean(dim=2), dim=1)
    else:
        L2 = torch.softmax(C, dim=1)                   # over atoms_2
        L1 = torch.softmax(C.transpose(1, 2), dim=1)   # (b, N1, N2), over atoms_1
        lt1, lt2 = a1 @ P["lt_layer_1/W"].T, a2 @ P["lt_layer_2/W"].T
        H1 = torch.tanh(lt1 + L1 @ lt2)
        H2 = torch.tanh(lt2 + L2 @ lt1)
        at1 = torch.softmax((H1 @ P["attention_layer_1/W"].T)[:, :, 0], dim=1)
        at2 = torch.softmax((H2 @ P["attention_layer_2/W"].T)[:, :, 0], dim=1)
    return (at1[:, :, None] * j1).sum(dim=1), (at2[:, :, None] * j2).sum(dim=1)


def circular_correlation(l, r):
    D = l.shape[1]
    idx = (torch.arange(D)[None, :] + torch.arange(D)[:, None]) % D      # [k, i] -> (i + k) mod D
    return torch.einsum("bi,bki->bk", l, r[:, idx])


def hole(P, l, r, n_hidden=0, act="relu", layers_name="hidden_layers"):
    h = circular_correlation(l, r)
    for i in range(n_hidden):
        h = ACT[act](lin(P, "%s/%d" % (layers_name, i), h))
    return lin(P, "l_out", h)


def sigmoid_cross_entropy(x, t):
    t = torch.as_tensor(t)
    keep = (t != -1)
    cnt = max(int(keep.sum()), 1)
    tt = t.to(x.dtype)
    per = torch.nn.functional.binary_cross_entropy_with_logits(x, tt.clamp(min=0), reduction="none")
    return (per * keep).sum() / cnt
