"""Fixtures produced by the REFERENCE'S OWN link files (tests/golden/make_reference_golden.py runs /root/reference/models/*.py,
unmodified, over oracle/chainer_shim): the oracle must reproduce them to 1e-10 in float64 (CPU), and the CUDA links must
reproduce them within the fp32 parity bound 1e-4 (GPU) -- outputs, input gradients and every parameter gradient."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import minichainer as M
from oracle import reference_path as R

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))
NAMES = [os.path.basename(f)[4:-4] for f in FIXTURES if not os.path.basename(f).startswith("ref_pair_")]
PAIRS = [os.path.basename(f)[9:-4] for f in FIXTURES if os.path.basename(f).startswith("ref_pair_")]


def _load(name):
    z = np.load(os.path.join(HERE, "golden", "ref_%s.npz" % name), allow_pickle=False)
    meta = eval(str(z["meta"]), {"__builtins__": {}}, {})

    def seq(prefix):
        out = []
        while "%s:%d" % (prefix, len(out)) in z.files:
            out.append(z["%s:%d" % (prefix, len(out))])
        return out
    return dict(meta=meta, params={k[6:]: z[k] for k in z.files if k.startswith("param:")}, ints=seq("int"), floats=seq("float"),
                ws=seq("w"), outs=seq("out"), gin={int(k[4:]): z[k] for k in z.files if k.startswith("gin:")},
                gparam={k[7:]: z[k] for k in z.files if k.startswith("gparam:")})


def _hidden(params, names):
    dims, i = [], 0
    while "%s/%d/W" % (names, i) in params:
        dims.append(params["%s/%d/W" % (names, i)].shape[0])
        i += 1
    return tuple(dims)


def _oracle_fn(fx, tab):
    """callable(*float Vars) -> tuple of output Vars, on the oracle"""
    m, p, P = fx["meta"], fx["params"], R.P(tab)
    kind = m["kind"]
    if kind == "ggnn":
        net = R.GGNN(P, m["O"], m["H"], m["T"], concat_hidden=m["concat"], weight_tying=m["tied"], activation=m["act"])
        return lambda A: (net(fx["ints"][0], A),)
    if kind == "gin":
        net = R.GIN(P, m["O"], m["H"], m["T"], concat_hidden=m["concat"], weight_tying=m["tied"], activation=m["act"])
        return lambda: (net(fx["ints"][0], fx["floats"][0]),)
    if kind == "nfp":
        net = R.NFP(P, m["O"], m["H"], m["T"])

        def fn():
            g = net(fx["ints"][0], fx["floats"][0])
            return g, net.get_atom_array()
        return fn
    if kind == "mono":
        net = R.GGNNMono(P, m["O"], m["H"], m["T"], weight_tying=m["tied"], sum_readout=m["sum_readout"])

        def fn(A):
            g = net(fx["ints"][0], A)
            return (g, net.get_atom_array()) if m["with_atoms"] else (g,)
        return fn
    if kind == "ggnn_update":
        link = R.GGNNUpdate(P, m["H"])

        def fn(h, A):
            link.reset_state()
            return (link(link(h, A), A),)
        return fn
    if kind == "relgcn":
        net = R.RelGCN(P, m["O"], ch_list=list(m["ch"]), scale_adj=m["scale"])
        return lambda: (net(fx["ints"][0], fx["floats"][0]),)
    if kind in ("coattn_alter", "coattn_para", "coattn_circ", "coattn_global", "coattn_neural"):
        cls = {"coattn_global": lambda: R.GlobalCoattention(P, m["H"], m["O"]),
               "coattn_neural": lambda: R.NeuralCoattention(P, m["H"], m["O"], activation="tanh"),
               "coattn_alter": lambda: R.AlternatingCoattention(P, m["H"], m["O"], m["head"]),
               "coattn_para": lambda: R.ParallelCoattention(P, m["H"], m["O"], m["head"]),
               "coattn_circ": lambda: R.CircularParallelCoattention(P, m["H"], m["O"])}[kind]()
        return lambda a1, a2, g1, g2: cls(a1, g1, a2, g2)
    if kind == "coattn_bimpm":
        net = R.BiMPM(P, m["H"], m["O"], m["head"])
        return lambda a1, a2: net(a1, None, a2, None)
    if kind.startswith("coattn"):
        cls = {"coattn_nie": lambda: R.NieFineCoattention(P, m["H"], m["O"], m["head"], activation="tanh"),
               "coattn_vqa": lambda: R.VQAParallelCoattention(P, m["H"], m["O"], m["head"]),
               "coattn_pool": lambda: R.PoolingFineCoattention(P, m["H"], m["O"]),
               "coattn_fourier": lambda: R.FourierFineCoattention(P, m["H"], m["O"], m["head"], activation="tanh"),
               "coattn_deep": lambda: R.DeepNieFineCoattention(P, m["H"], m["O"], m["head"], activation="tanh"),
               "coattn_very_deep": lambda: R.VeryDeepNieFineCoattention(P, m["H"], m["O"], m["head"], activation="tanh"),
               "coattn_extreme_deep": lambda: R.ExtremeDeepNieFineCoattention(P, m["H"], m["O"], m["head"], activation="tanh")}[kind]()
        return lambda a1, a2: cls(a1, None, a2, None)
    if kind == "readout":
        net = R.GGNNReadout(P, m["O"], m["H"], nobias=m["nobias"], activation=m["act"], activation_agg=m["agg"])
        mask = fx["floats"][2] if m["use_mask"] else None
        return lambda h, h0: (net(h, h0 if m["use_h0"] else None, mask),)
    D, K = m["D"], m["K"]
    if kind == "head_hole":
        net = R.HolE(P, K, hidden_dims=_hidden(p, "hidden_layers"))
    elif kind == "head_hole_mlp_py":
        net = R.HolE(P, K, hidden_dims=_hidden(p, "layers"), layers_name="layers")
    elif kind == "head_symmlp":
        net = R.SymMLP(P, K, _hidden(p, "layers"))
    elif kind == "head_ntn":
        net = R.NTN(P, D, D, K, p["ntn_layer/b"].shape[0], _hidden(p, "mlp_layers"))
    elif kind == "head_distmult":
        net = R.DistMult(P, D, D, K, p["dm_layer/W"].shape[0], _hidden(p, "mlp_layers"))
    else:
        mlp = R.MLP(P, K, _hidden(p, "layers"))
        return lambda l, r: (mlp(M.concat((l, r), axis=-1)),)
    return lambda l, r: (net(l, r),)


def _n_var_inputs(fx):
    k = fx["meta"]["kind"]
    return {"ggnn": 1, "mono": 1, "ggnn_update": 2, "relgcn": 0, "gin": 0, "nfp": 0, "readout": 2, "coattn_alter": 4, "coattn_para": 4, "coattn_circ": 4, "coattn_global": 4, "coattn_neural": 4}.get(k, 2)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference_generated_fixture(name):
    fx = _load(name)
    tab = R.wrap_params(fx["params"])
    vs = [M.param(np.array(x)) for x in fx["floats"][:_n_var_inputs(fx)]]
    outs = _oracle_fn(fx, tab)(*vs)
    total = None
    for o, w in zip(outs, fx["ws"]):
        t = M.sum_(M.mul(o, M.const(w)))
        total = t if total is None else M.add(total, t)
    total.backward()
    assert len(outs) == len(fx["outs"])
    for o, ref in zip(outs, fx["outs"]):
        np.testing.assert_allclose(o.data, ref, rtol=1e-10, atol=1e-12)
    for i, g in fx["gin"].items():
        np.testing.assert_allclose(vs[i].grad, g, rtol=1e-9, atol=1e-12)
    assert fx["gparam"], "fixture without parameter gradients"
    for k, g in fx["gparam"].items():
        np.testing.assert_allclose(tab[k].grad, g, rtol=1e-9, atol=1e-12, err_msg=k)


def test_fixture_set_covers_every_hot_path_file():
    kinds = {_load(n)["meta"]["kind"] for n in NAMES}
    assert {"ggnn", "mono", "ggnn_update", "relgcn", "coattn_nie", "coattn_vqa", "coattn_pool", "readout", "head_hole",
            "head_hole_mlp_py", "head_mlp", "head_symmlp", "head_ntn", "head_distmult", "nfp", "coattn_bimpm"} <= kinds


def _pair_fixture(name):
    z = np.load(os.path.join(HERE, "golden", "ref_pair_%s.npz" % name), allow_pickle=False)
    meta = eval(str(z["meta"]), {"__builtins__": {}}, {})
    return meta, z["logits"], float(z["loss"]), {k[7:]: z[k] for k in z.files if k.startswith("gparam:")}


@pytest.mark.parametrize("name", PAIRS)
def test_oracle_reproduces_reference_pair_step(name):
    """The whole unit of the metric -- the reference's GraphConvPredictorForPair (train_binary.py:59-141 with co-attention,
    train_ddi_modify_eval2.py:50-104 without) over its own encoder / attention / HolE files + sigmoid-CE, fwd+bwd."""
    import cases
    meta, logits, loss, gp = _pair_fixture(name)
    o = cases.oracle_eval(cases.pair_case(meta["case"], seed=meta["seed"]))
    np.testing.assert_allclose(o["logits"], logits, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(float(o["loss"]), loss, rtol=1e-10)
    assert len(gp) >= 10
    for k, g in gp.items():
        np.testing.assert_allclose(o["grads"][k], g, rtol=1e-9, atol=1e-13, err_msg=k)


def test_pair_fixtures_cover_the_baseline_shapes():
    # CB: the bench's exact shape; L1: 100 / 70 atoms per molecule (no atom cap in the reference); E3 / E3B: [h_first | h_last] atoms
    assert set(PAIRS) >= {"A", "B", "C", "U", "MU", "CB", "L1", "E3", "E3B"}


@pytest.mark.gpu
@pytest.mark.parametrize("name", PAIRS)
def test_cuda_pair_step_reproduces_reference_pair_step(name):
    import cases
    import product
    from product import rel_err
    meta, logits, loss, gp = _pair_fixture(name)
    p = product.product_eval(cases.pair_case(meta["case"], seed=meta["seed"]))
    assert rel_err(p["logits"], logits) <= 1e-4
    assert abs(float(p["loss"]) - loss) <= 1e-4 * max(1.0, abs(loss))
    for k, g in gp.items():
        if np.abs(g).max() > 1e-12:
            # a 1-element gradient that is a cancelling sum over every atom pair (the energy bias: ~2e4 terms at 100 x 70 atoms, result
            # ~1e-4) carries fp32 summation noise of ~2e-8 absolute: held to 1e-7 absolute instead of 1e-4 of its own magnitude
            assert rel_err(p["grads"][k], g, floor=1e-3 if g.size == 1 else 1e-30) <= 1e-4, k


# ------------------------------------------------------------------------------------------------ GPU
def _product(fx):
    """(callable(*float tensors) -> tuple of output tensors, link) on the CUDA links"""
    import gcnbmp
    f = gcnbmp.functions
    m, p = fx["meta"], fx["params"]
    kind = m["kind"]
    act = lambda a: getattr(f, a)
    if kind == "ggnn":
        net = gcnbmp.GGNN(m["O"], hidden_dim=m["H"], n_layers=m["T"], concat_hidden=m["concat"], weight_tying=m["tied"], activation=act(m["act"]))
        return (lambda A: (net(fx["ints"][0], A),)), net
    if kind == "gin":
        net = gcnbmp.GIN(m["O"], hidden_dim=m["H"], n_layers=m["T"], dropout_ratio=0.0, concat_hidden=m["concat"], weight_tying=m["tied"], activation=act(m["act"]))
        return (lambda: (net(fx["ints"][0], fx["floats"][0].astype(np.float32)),)), net
    if kind == "nfp":
        net = gcnbmp.NFP(m["O"], hidden_dim=m["H"], n_layers=m["T"])

        def fn():
            g = net(fx["ints"][0], fx["floats"][0].astype(np.float32))
            return g, net.get_atom_array()
        return fn, net
    if kind == "mono":
        net = gcnbmp.GGNNMono(m["O"], m["H"], m["T"], weight_tying=m["tied"], sum_readout=m["sum_readout"])

        def fn(A):
            g = net(fx["ints"][0], A)
            return (g, net.get_atom_array()) if m["with_atoms"] else (g,)
        return fn, net
    if kind == "ggnn_update":
        link = gcnbmp.GGNNUpdate(m["H"])

        def fn(h, A):
            link.reset_state()
            return (link(link(h, A), A),)
        return fn, link
    if kind == "relgcn":
        net = gcnbmp.RelGCN(m["O"], ch_list=list(m["ch"]), scale_adj=m["scale"])
        return (lambda: (net(fx["ints"][0], fx["floats"][0].astype(np.float32)),)), net
    if kind in ("coattn_alter", "coattn_para", "coattn_circ", "coattn_global", "coattn_neural"):
        net = {"coattn_global": lambda: gcnbmp.GlobalCoattention(m["H"], m["O"]),
               "coattn_neural": lambda: gcnbmp.NeuralCoattention(m["H"], m["O"], activation=f.tanh),
               "coattn_alter": lambda: gcnbmp.AlternatingCoattention(m["H"], m["O"], m["head"]),
               "coattn_para": lambda: gcnbmp.ParallelCoattention(m["H"], m["O"], m["head"]),
               "coattn_circ": lambda: gcnbmp.CircularParallelCoattention(m["H"], m["O"])}[kind]()
        return (lambda a1, a2, g1, g2: net(a1, g1, a2, g2)), net
    if kind == "coattn_bimpm":
        net = gcnbmp.BiMPM(m["H"], m["O"], m["head"])
        return (lambda a1, a2: net(a1, None, a2, None)), net
    if kind.startswith("coattn"):
        net = {"coattn_nie": lambda: gcnbmp.NieFineCoattention(m["H"], m["O"], m["head"], activation=f.tanh),
               "coattn_vqa": lambda: gcnbmp.VQAParallelCoattention(m["H"], m["O"], m["head"]),
               "coattn_pool": lambda: gcnbmp.PoolingFineCoattention(m["H"], m["O"]),
               "coattn_fourier": lambda: gcnbmp.FourierFineCoattention(m["H"], m["O"], m["head"], activation=f.tanh),
               "coattn_deep": lambda: gcnbmp.DeepNieFineCoattention(m["H"], m["O"], m["head"], activation=f.tanh),
               "coattn_very_deep": lambda: gcnbmp.VeryDeepNieFineCoattention(m["H"], m["O"], m["head"], activation=f.tanh),
               "coattn_extreme_deep": lambda: gcnbmp.ExtremeDeepNieFineCoattention(m["H"], m["O"], m["head"], activation=f.tanh)}[kind]()
        return (lambda a1, a2: net(a1, None, a2, None)), net
    if kind == "readout":
        net = gcnbmp.GGNNReadout(m["O"], m["H"], nobias=m["nobias"], activation=act(m["act"]), activation_agg=act(m["agg"]))
        mask = fx["floats"][2].astype(np.float32) if m["use_mask"] else None
        return (lambda h, h0: (net(h, h0 if m["use_h0"] else None, mask),)), net
    D, K = m["D"], m["K"]
    if kind == "head_hole":
        net = gcnbmp.HolE(K, hidden_dims=_hidden(p, "hidden_layers"))
    elif kind == "head_hole_mlp_py":
        net = gcnbmp.HolE(K, hidden_dims=_hidden(p, "layers"), layers_name="layers")
    elif kind == "head_symmlp":
        net = gcnbmp.SymMLP(K, _hidden(p, "layers"))
    elif kind == "head_ntn":
        net = gcnbmp.NTN(D, D, K, p["ntn_layer/b"].shape[0], _hidden(p, "mlp_layers"))
    elif kind == "head_distmult":
        net = gcnbmp.DistMult(D, D, K, p["dm_layer/W"].shape[0], _hidden(p, "mlp_layers"))
    else:
        net = gcnbmp.MLP(K, _hidden(p, "layers"))
        return (lambda l, r: (net(gcnbmp.functional.PairFeatures.apply(l, r, gcnbmp._capi.PAIR_CONCAT)),)), net
    return (lambda l, r: (net(l, r),)), net


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_links_reproduce_reference_generated_fixture(name):
    from product import rel_err
    fx = _load(name)
    fn, link = _product(fx)
    link.load_params(fx["params"])
    kind = fx["meta"]["kind"]
    vs = []
    for i, x in enumerate(fx["floats"][:_n_var_inputs(fx)]):
        is_adj = x.ndim == 4
        vs.append(torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=not is_adj))
    outs = fn(*vs)
    assert len(outs) == len(fx["outs"])
    total = sum((o * torch.tensor(w, dtype=torch.float32, device="cuda")).sum() for o, w in zip(outs, fx["ws"]))
    total.backward()
    for o, ref in zip(outs, fx["outs"]):
        assert rel_err(o.detach().cpu().numpy(), ref) <= 1e-4
    for i, g in fx["gin"].items():
        if vs[i].requires_grad:
            assert rel_err(vs[i].grad.cpu().numpy(), g) <= 1e-4, "input gradient %d" % i
    got = link.grad_dict()
    for k, g in fx["gparam"].items():
        if np.abs(g).max() > 1e-12:
            assert rel_err(got[k], g) <= 1e-4, k
