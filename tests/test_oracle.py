"""CPU tests (run in the build container): pin the NumPy oracle with an independent Torch
autograd twin, finite differences and algebraic identities.  The reference has no tests or
golden vectors for this path and Chainer cannot run here, so these pins
are what stands behind the oracle -- "parity unpinned" in the sense of the task statement."""
import numpy as np
import pytest
import torch

import cases
import torch_twin as TT
from oracle import minichainer as F
from oracle import reference_path as R


def twin_eval(case):
    sp = case["spec"]
    P = TT.T(case["params"])
    a1, A1, a2, A2 = case["inputs"]
    A1, A2 = torch.tensor(A1), torch.tensor(A2)
    sub = lambda pre: {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}
    G = sub("graph_conv/")

    def enc(a, A):
        if sp["enc"] == "mono":
            return TT.ggnn_mono(G, a, A, sp["T"], sp["tied"], sp["sum_readout"])
        if sp["enc"] == "ggnn":
            return TT.ggnn(G, a, A, sp["T"], sp["tied"], activation=sp.get("activation", "identity"))
        return TT.relgcn(G, a, A, sp["ch"], sp["scale_adj"])
    g1, x1 = enc(a1, A1)
    g2, x2 = enc(a2, A2)
    if sp["attn"]:
        g1, g2 = TT.coattention(sub("attn/"), x1, x2, "pool" if sp["attn"] == "pool" else "fine", "tanh")
    logits = TT.hole(sub("mlp/"), g1, g2, len(sp["hole_hidden"]))
    loss = TT.sigmoid_cross_entropy(logits, case["labels"])
    loss.backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in P.items()}
    return dict(loss=loss.detach().numpy(), logits=logits.detach().numpy(), grads=grads)


@pytest.mark.parametrize("name", ["A", "C", "U", "M", "MU", "B"])
def test_oracle_matches_torch_twin_fp64(name):
    case = cases.pair_case(name, seed=11)
    o = cases.oracle_eval(case)
    t = twin_eval(case)
    np.testing.assert_allclose(o["logits"], t["logits"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(o["loss"], t["loss"], rtol=1e-10)
    for k in o["grads"]:
        scale = max(np.abs(t["grads"][k]).max(), 1e-12)
        np.testing.assert_allclose(o["grads"][k] / scale, t["grads"][k] / scale, rtol=0, atol=1e-9, err_msg=k)


def test_oracle_finite_differences():
    case = cases.pair_case("U", seed=3)
    o = cases.oracle_eval(case)
    rng = np.random.default_rng(0)
    eps = 1e-6
    for key in ["graph_conv/update_layer/U/W", "graph_conv/message_layers/1/b", "attn/energy_layer/W",
                "attn/lt_layer_2/W", "mlp/l_out/W", "graph_conv/embed/W", "graph_conv/i_layers/0/W"]:
        base = case["params"][key]
        idxs = [tuple(rng.integers(0, s) for s in base.shape) for _ in range(3)]
        if key.endswith("embed/W"):
            idxs = [(6, 1), (0, 3), (8, 0)]      # carbon, padding id 0 (a live row!), oxygen
        for idx in idxs:
            vals = []
            for sgn in (+1, -1):
                p2 = dict(case["params"])
                w = base.copy()
                w[idx] += sgn * eps
                p2[key] = w
                vals.append(float(cases.oracle_eval(dict(case, params=p2))["loss"]))
            fd = (vals[0] - vals[1]) / (2 * eps)
            assert abs(fd - o["grads"][key][idx]) <= 1e-6 * max(1.0, abs(fd)), (key, idx, fd, o["grads"][key][idx])


def test_padding_atoms_are_live():
    """Reference quirk: id-0 atoms get a real embedding row, evolve through the GRU
    and are summed by the readout -> the embedding row 0 receives gradient."""
    case = cases.pair_case("A", seed=5)
    o = cases.oracle_eval(case)
    assert np.abs(o["grads"]["graph_conv/embed/W"][0]).max() > 0


def test_tied_ggnn_equals_threaded_single_steps():
    """GGNN driver with tied weights == calling ONE GGNNUpdate link T times (state threaded)."""
    rng = np.random.default_rng(2)
    H, T, mb, N = 8, 3, 2, 7
    shapes = R.ggnn_shapes(4, H, T)
    params = R.init_params(shapes, rng, dtype=np.float64)
    from gcnbmp import synthetic
    atoms, adj = synthetic.random_molecules(rng, mb, N)
    adj = adj.astype(np.float64)
    tab = R.wrap_params(params)
    net = R.GGNN(R.P(tab), 4, H, T)
    net(atoms, adj)
    full = net.get_atom_array().data
    upd = R.GGNNUpdate(R.P(tab).sub("update_layers/0"), H)
    upd.reset_state()
    h = F.embed_id(atoms, tab["embed/W"])
    for _ in range(T):
        h = upd(h, adj)
    np.testing.assert_allclose(h.data, full, rtol=1e-13)


def test_message_interleave_identity():
    """y[b,n,c*E+e] -> M[b,e,n,c] (ggnn_update.py:35-39) == per-edge weights W_e = W[e::E]."""
    rng = np.random.default_rng(4)
    H, E, mb, N = 6, 4, 3, 5
    W, b = rng.standard_normal((E * H, H)), rng.standard_normal(E * H)
    h, adj = rng.standard_normal((mb, N, H)), rng.random((mb, E, N, N))
    upd = R.GGNNUpdate(R.P({"graph_linear/W": F.const(W), "graph_linear/b": F.const(b)}), H, E)
    m = upd.message(F.const(h), adj).data
    ref = sum(adj[:, e] @ (h @ W[e::E].T + b[e::E]) for e in range(E))
    np.testing.assert_allclose(m, ref, rtol=1e-13)


def test_hole_fft_equals_direct_sum():
    rng = np.random.default_rng(6)
    l, r = rng.standard_normal((4, 16)), rng.standard_normal((4, 16))
    hole = R.HolE(R.P({}), 1, ())
    c = hole.circular_correlation(F.const(l), F.const(r)).data
    direct = np.stack([[sum(l[b, i] * r[b, (i + k) % 16] for i in range(16)) for k in range(16)] for b in range(4)])
    np.testing.assert_allclose(c, direct, rtol=1e-12, atol=1e-13)


def test_relgcn_rejects_bad_input_type():
    with pytest.raises(ValueError):
        R.RelGCN(R.P({}), input_type="complex")


def test_fp32_oracle_close_to_fp64():
    """The fp32 'reference CPU path' stays within the 1e-4 budget of the fp64 ground truth."""
    case = cases.pair_case("C", seed=9)
    o64 = cases.oracle_eval(case)
    case32 = dict(case, params={k: v.astype(np.float32) for k, v in case["params"].items()})
    o32 = cases.oracle_eval(case32, dtype=np.float32)
    err = np.abs(o32["logits"] - o64["logits"]).max() / np.abs(o64["logits"]).max()
    assert err < 1e-4
