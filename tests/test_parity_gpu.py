"""GPU parity tests (run on the B200 box): the CUDA path, called through the drop-in links ->
ctypes C-ABI, against the fp64 NumPy oracle on the same seeded inputs.
Bar (BASELINE.json north_star): <= 1e-4 relative in fp32 mode, forward AND gradients."""
import numpy as np
import pytest
import torch

import cases
import product
from product import rel_err
from oracle import minichainer as F
from oracle import reference_path as R

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("name", ["A", "C", "U", "M", "MU", "B"])
def test_pair_forward_backward_matches_oracle(name):
    case = cases.pair_case(name, seed=11)
    o = cases.oracle_eval(case)
    p = product.product_eval(case)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    assert abs(p["loss"] - float(o["loss"])) <= TOL * max(1.0, abs(float(o["loss"])))
    assert set(p["grads"]) == set(o["grads"])
    for k in sorted(o["grads"]):
        assert p["grads"][k].shape == o["grads"][k].shape, k
        assert rel_err(p["grads"][k], o["grads"][k]) <= TOL, (k, rel_err(p["grads"][k], o["grads"][k]))


@pytest.mark.parametrize("name", ["L1", "L2"])
def test_pair_with_more_than_64_atoms_matches_oracle(name):
    """N in {65, 70, 100, 128}: the reference has no atom cap (ggnn_preprocessor.py:41, max_atoms=-1; concat_mols pads to the batch
    maximum).  The GGNN encoder runs these on the fp32 tensor-core path (its row GEMMs never see molecule boundaries), read-out and
    co-attention as library-GEMM compositions of the same formulas; logits, loss and every parameter gradient vs the fp64 oracle."""
    case = cases.pair_case(name, seed=5)
    assert max(case["spec"]["N1"], case["spec"]["N2"]) > 64
    o = cases.oracle_eval(case)
    p = product.product_eval(case)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    assert abs(p["loss"] - float(o["loss"])) <= TOL * max(1.0, abs(float(o["loss"])))
    assert set(p["grads"]) == set(o["grads"])
    for k in sorted(o["grads"]):
        # a 1-element gradient that is a cancelling sum over all atom pairs (the energy bias) is held to 1e-7 absolute
        err = rel_err(p["grads"][k], o["grads"][k], floor=1e-3 if o["grads"][k].size == 1 else 1e-30)
        assert err <= TOL, (k, err)
    with torch.no_grad():
        model = product.product_model(case["spec"], case["params"])
        a1, A1, a2, A2 = case["inputs"]
        y_eval = model(a1, A1.astype(np.float32), a2, A2.astype(np.float32)).cpu().numpy()
    assert rel_err(y_eval, o["logits"]) <= TOL


def test_coattention_on_256_wide_atoms_and_readout_backward_at_hidden_256():
    """train_ddi_modify_eval3.py:110-134 at GGNN hidden 128 feeds [h_first | h_last] = 256-wide atoms to the co-attention (wider than the
    shared-memory-resident CUDA kernel holds), and BASELINE config D's read-out (hidden = out_dim = 256) as a training step: both take
    the library-GEMM composition of the same formulas; every value and gradient vs the fp64 oracle."""
    case = cases.pair_case("E3B", seed=8)
    o = cases.oracle_eval(case)
    p = product.product_eval(case)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    for k in sorted(o["grads"]):
        err = rel_err(p["grads"][k], o["grads"][k], floor=1e-3 if o["grads"][k].size == 1 else 1e-30)
        assert err <= TOL, (k, err)
    import gcnbmp
    rng = np.random.default_rng(3)
    mb, N, H, O = 3, 50, 256, 256
    shapes = {"i_layer/W": (O, 2 * H), "i_layer/b": (O,), "j_layer/W": (O, 2 * H), "j_layer/b": (O,)}
    params = R.init_params(shapes, rng, dtype=np.float64)
    tab = R.wrap_params(params)
    h, h0, w = rng.standard_normal((mb, N, H)) * 0.3, rng.standard_normal((mb, N, H)) * 0.3, rng.standard_normal((mb, O))
    vh, vh0 = F.param(h), F.param(h0)
    og = R.GGNNReadout(R.P(tab), O, H, activation="tanh", activation_agg="tanh")(vh, vh0)
    F.sum_(F.mul(og, F.const(w))).backward()
    net = gcnbmp.GGNNReadout(O, H, activation=gcnbmp.functions.tanh, activation_agg=gcnbmp.functions.tanh)
    net.load_params(params)
    th = torch.tensor(h, dtype=torch.float32, device="cuda", requires_grad=True)
    th0 = torch.tensor(h0, dtype=torch.float32, device="cuda", requires_grad=True)
    pg = net(th, th0)
    (pg * torch.tensor(w, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(pg.detach().cpu().numpy(), og.data) <= TOL
    assert rel_err(th.grad.cpu().numpy(), vh.grad) <= TOL and rel_err(th0.grad.cpu().numpy(), vh0.grad) <= TOL
    g = net.grad_dict()
    for k in params:
        assert rel_err(g[k], tab[k].grad) <= TOL, k


def test_bf16_mode_takes_batches_with_more_than_64_atoms_through_the_fp32_tensor_core_path():
    """A model in BF16 mode does not fail on a batch padded beyond 64 atoms: that batch runs the fp32 tensor-core encoder and the
    library compositions, i.e. inside the fp32 bound."""
    import gcnbmp
    case = cases.pair_case("L1", seed=5)
    o = cases.oracle_eval(case)
    p = product.product_eval(case, mode=gcnbmp.MODE_BF16)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    for k in sorted(o["grads"]):
        assert rel_err(p["grads"][k], o["grads"][k], floor=1e-3 if o["grads"][k].size == 1 else 1e-30) <= TOL, k
    a1, A1, a2, A2 = case["inputs"]
    model = product.product_model(case["spec"], case["params"])
    model.graph_conv.mode = model.attn.mode = gcnbmp.MODE_BF16
    with torch.no_grad():           # uint8 adjacency as the BF16-mode callers hold it
        y8 = model(a1, A1.astype(np.uint8), a2, A2.astype(np.uint8)).cpu().numpy()
    assert rel_err(y8, o["logits"]) <= TOL


def test_more_than_64_atoms_outside_the_tensor_core_shapes_fails_loudly():
    import gcnbmp
    from gcnbmp import synthetic
    atoms, adj = synthetic.random_molecules(np.random.default_rng(0), 3, 80)
    net = gcnbmp.GGNNMono(16, 32, 2)            # hidden 32: only the <= 64-atom FFMA kernels cover it
    with pytest.raises(ValueError, match="n_atoms=80"):
        net(atoms, adj)


def test_inference_path_equals_training_path():
    """no_grad takes the stash-free kernel path; outputs must be identical bit for bit."""
    case = cases.pair_case("C", seed=2)
    model = product.product_model(case["spec"], case["params"])
    a1, A1, a2, A2 = case["inputs"]
    A1, A2 = A1.astype(np.float32), A2.astype(np.float32)
    y_train = model(a1, A1, a2, A2).detach().cpu().numpy()
    with torch.no_grad():
        y_eval = model(a1, A1, a2, A2).cpu().numpy()
    np.testing.assert_array_equal(y_train, y_eval)


def _np_params(shapes, seed):
    return R.init_params(shapes, np.random.default_rng(seed), dtype=np.float64)


def test_ggnn_update_link_state_threading():
    """GGNNUpdate called repeatedly keeps the StatefulGRU state; after reset_state the first
    call takes the stateless branch (models/update/ggnn_update.py:31-66)."""
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(1)
    H, mb, N = 32, 3, 19
    shapes = {k[len("update_layers/0/"):]: v for k, v in R.ggnn_shapes(8, H, 1).items() if k.startswith("update_layers/0/")}
    params = _np_params(shapes, 5)
    _, adj = synthetic.random_molecules(rng, mb, N)
    h0 = rng.standard_normal((mb, N, H))
    other = rng.standard_normal((mb, N, H))
    # oracle: 3 calls, the third with an input that is NOT the state
    tab = R.wrap_params(params)
    hv, ov = F.param(h0), F.param(other)
    upd = R.GGNNUpdate(R.P(tab), H)
    o1 = upd(hv, adj.astype(np.float64))
    o2 = upd(o1, adj.astype(np.float64))
    o3 = upd(ov, adj.astype(np.float64))          # state (o2) != input (other)
    F.sum_(F.mul(o3, o3)).backward()
    link = gcnbmp.GGNNUpdate(hidden_dim=H)
    link.load_params(params)
    ht = torch.tensor(h0, dtype=torch.float32, device="cuda", requires_grad=True)
    ot = torch.tensor(other, dtype=torch.float32, device="cuda", requires_grad=True)
    A = torch.tensor(adj, device="cuda")
    p1 = link(ht, A)
    p2 = link(p1, A)
    p3 = link(ot, A)
    (p3 * p3).sum().backward()
    assert rel_err(p1.detach().cpu().numpy(), o1.data) <= TOL
    assert rel_err(p3.detach().cpu().numpy(), o3.data) <= TOL
    assert rel_err(ht.grad.cpu().numpy(), hv.grad) <= TOL
    assert rel_err(ot.grad.cpu().numpy(), ov.grad) <= TOL
    g = link.grad_dict()
    for k in g:
        assert rel_err(g[k], tab[k].grad) <= TOL, k
    link.reset_state()
    p4 = link(ht.detach(), A)
    assert rel_err(p4.detach().cpu().numpy(), o1.data) <= TOL


@pytest.mark.parametrize("act,agg,use_h0,use_mask,nobias", [
    ("identity", "identity", True, False, False), ("tanh", "tanh", True, True, False),
    ("tanh", "identity", False, False, True), ("relu", "sigmoid", True, True, False)])
def test_readout_variants(act, agg, use_h0, use_mask, nobias):
    import gcnbmp
    rng = np.random.default_rng(3)
    mb, N, H, O = 4, 21, 32, 24
    kin = 2 * H if use_h0 else H
    shapes = {"i_layer/W": (O, kin), "j_layer/W": (O, kin)}
    if not nobias:
        shapes.update({"i_layer/b": (O,), "j_layer/b": (O,)})
    params = _np_params(shapes, 7)
    h, h0 = rng.standard_normal((mb, N, H)), rng.standard_normal((mb, N, H))
    mask = (rng.random((mb, N)) < 0.7).astype(np.float64) if use_mask else None
    tab = R.wrap_params(params)
    hv, h0v = F.param(h), F.param(h0)
    ro = R.GGNNReadout(R.P(tab), O, H, nobias=nobias, activation=act, activation_agg=agg)
    og = ro(hv, h0v if use_h0 else None, mask)
    w = rng.standard_normal(og.shape)
    F.sum_(F.mul(og, F.const(w))).backward()
    f = gcnbmp.functions
    link = gcnbmp.GGNNReadout(O, H, nobias=nobias, activation=getattr(f, act), activation_agg=getattr(f, agg))
    link.load_params(params)
    ht = torch.tensor(h, dtype=torch.float32, device="cuda", requires_grad=True)
    h0t = torch.tensor(h0, dtype=torch.float32, device="cuda", requires_grad=True)
    pg = link(ht, h0t if use_h0 else None, None if mask is None else mask.astype(np.float32))
    (pg * torch.tensor(w, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(pg.detach().cpu().numpy(), og.data) <= TOL
    assert rel_err(ht.grad.cpu().numpy(), hv.grad) <= TOL
    if use_h0:
        assert rel_err(h0t.grad.cpu().numpy(), h0v.grad) <= TOL
    g = link.grad_dict()
    for k in g:
        assert rel_err(g[k], tab[k].grad) <= TOL, k


def test_relgcn_update_bare_link_and_float_input():
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(8)
    mb, N, cin, cout = 3, 17, 16, 40
    _, adj = synthetic.random_molecules(rng, mb, N)
    adj = adj * rng.random(adj.shape).astype(np.float32)      # fractional adjacency
    h = rng.standard_normal((mb, N, cin))
    shapes = {"graph_linear_self/W": (cout, cin), "graph_linear_self/b": (cout,),
              "graph_linear_edge/W": (cout * 4, cin), "graph_linear_edge/b": (cout * 4,)}
    params = _np_params(shapes, 2)
    tab = R.wrap_params(params)
    hv = F.param(h)
    o = R.RelGCNUpdate(R.P(tab), cin, cout)(hv, adj.astype(np.float64))
    F.sum_(F.mul(o, o)).backward()
    link = gcnbmp.RelGCNUpdate(cin, cout)
    link.load_params(params)
    ht = torch.tensor(h, dtype=torch.float32, device="cuda", requires_grad=True)
    p = link(ht, adj)
    (p * p).sum().backward()
    assert rel_err(p.detach().cpu().numpy(), o.data) <= TOL
    assert rel_err(ht.grad.cpu().numpy(), hv.grad) <= TOL
    g = link.grad_dict()
    for k in g:
        assert rel_err(g[k], tab[k].grad) <= TOL, k


@pytest.mark.parametrize("variant,n1,n2,H,O", [("nie", 64, 64, 128, 128), ("vqa", 13, 37, 32, 16),
                                                ("pool", 50, 9, 64, 32), ("nie", 1, 5, 16, 8)])
def test_coattention_ragged(variant, n1, n2, H, O):
    import gcnbmp
    rng = np.random.default_rng(4)
    mb, head = 3, 8
    params = _np_params(R.coattn_shapes(H, O, None if variant == "pool" else head), 6)
    a1, a2 = rng.standard_normal((mb, n1, H)) * 0.5, rng.standard_normal((mb, n2, H)) * 0.5
    tab = R.wrap_params(params)
    v1, v2 = F.param(a1), F.param(a2)
    if variant == "pool":
        oc = R.PoolingFineCoattention(R.P(tab), H, O)
        link = gcnbmp.PoolingFineCoattention(H, O)
    elif variant == "vqa":
        oc = R.VQAParallelCoattention(R.P(tab), H, O, head)
        link = gcnbmp.VQAParallelCoattention(H, O, head)
    else:
        oc = R.NieFineCoattention(R.P(tab), H, O, head)       # identity activation default
        link = gcnbmp.NieFineCoattention(H, O, head)
    c1, c2 = oc(v1, None, v2, None)
    w1, w2 = rng.standard_normal(c1.shape), rng.standard_normal(c2.shape)
    F.add(F.sum_(F.mul(c1, F.const(w1))), F.sum_(F.mul(c2, F.const(w2)))).backward()
    link.load_params(params)
    t1 = torch.tensor(a1, dtype=torch.float32, device="cuda", requires_grad=True)
    t2 = torch.tensor(a2, dtype=torch.float32, device="cuda", requires_grad=True)
    p1, p2 = link(t1, None, t2, None)
    dev = lambda x: torch.tensor(x, dtype=torch.float32, device="cuda")
    ((p1 * dev(w1)).sum() + (p2 * dev(w2)).sum()).backward()
    assert rel_err(p1.detach().cpu().numpy(), c1.data) <= TOL
    assert rel_err(p2.detach().cpu().numpy(), c2.data) <= TOL
    assert rel_err(t1.grad.cpu().numpy(), v1.grad) <= TOL
    assert rel_err(t2.grad.cpu().numpy(), v2.grad) <= TOL
    g = link.grad_dict()
    for k in g:
        assert rel_err(g[k], tab[k].grad, floor=1e-3) <= TOL, k


@pytest.mark.parametrize("D", [16, 33, 256])
def test_hole_correlation_matches_fft_oracle(D):
    import gcnbmp
    rng = np.random.default_rng(5)
    l, r = rng.standard_normal((37, D)), rng.standard_normal((37, D))
    lv, rv = F.param(l), F.param(r)
    oc = R.HolE(R.P({}), 1, ()).circular_correlation(lv, rv)
    w = rng.standard_normal(oc.shape)
    F.sum_(F.mul(oc, F.const(w))).backward()
    lt = torch.tensor(l, dtype=torch.float32, device="cuda", requires_grad=True)
    rt = torch.tensor(r, dtype=torch.float32, device="cuda", requires_grad=True)
    pc = gcnbmp.HolE(1, ()).circular_correlation(lt, rt)
    (pc * torch.tensor(w, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(pc.detach().cpu().numpy(), oc.data) <= TOL
    assert rel_err(lt.grad.cpu().numpy(), lv.grad) <= TOL
    assert rel_err(rt.grad.cpu().numpy(), rv.grad) <= TOL


def test_shape_errors_are_loud():
    import gcnbmp
    m = gcnbmp.GGNNMono(8, 16, 2)
    atoms = np.zeros((1, 65), np.int32)
    adj = np.zeros((1, 4, 65, 65), np.float32)
    with pytest.raises(ValueError):
        m(atoms, adj)
    with pytest.raises(ValueError):
        gcnbmp.RelGCN(input_type="complex")
    with pytest.raises(RuntimeError):
        gcnbmp.functional.HoleCorr.apply(torch.zeros(2, 4), torch.zeros(2, 4))   # CPU tensors: no fallback


def test_single_molecule_and_hidden16_default():
    """mb = 1, reference default hidden_dim = 16 (GGNN.__init__ default)."""
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(12)
    shapes = R.ggnn_shapes(16, 16, 4)
    params = _np_params(shapes, 13)
    atoms, adj = synthetic.random_molecules(rng, 1, 9)
    tab = R.wrap_params(params)
    og = R.GGNN(R.P(tab), 16, 16, 4)(atoms, adj.astype(np.float64))
    net = gcnbmp.GGNN(16)
    net.load_params(params)
    pg = net(atoms, adj)
    assert rel_err(pg.detach().cpu().numpy(), og.data) <= TOL


def test_persistent_loops_many_molecules_per_cta():
    """mb > #SMs: every kernel walks several molecules/pairs per CTA (grid-stride loops)."""
    case = cases.pair_case("U", seed=21)
    sp = dict(case["spec"], mb=333)
    from gcnbmp import synthetic
    rng = np.random.default_rng(77)
    a1, A1 = synthetic.random_molecules(rng, 333, sp["N1"])
    a2, A2 = synthetic.random_molecules(rng, 333, sp["N2"])
    y = (rng.random((333, 1)) < 0.33).astype(np.int32)
    big = dict(case, spec=sp, inputs=(a1, A1.astype(np.float64), a2, A2.astype(np.float64)), labels=y)
    o = cases.oracle_eval(big)
    p = product.product_eval(big)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    for k in sorted(o["grads"]):
        assert rel_err(p["grads"][k], o["grads"][k], floor=1e-6) <= TOL, k


def test_trainer_microbatching_and_flat_buffers():
    """PairTrainer: chunked accumulation over flat parameter/gradient buffers (the layout the NCCL
    allreduce and Adam use) reproduces the single-shot gradient; host inputs stream correctly."""
    from gcnbmp.train import PairTrainer
    case = cases.pair_case("C", seed=31)
    o = cases.oracle_eval(case)
    model = product.product_model(case["spec"], case["params"])
    tr = PairTrainer(model, chunk=3, optimizer=False)
    a1, A1, a2, A2 = case["inputs"]
    y = case["labels"]
    count = float((y != -1).sum())
    host = [torch.from_numpy(np.ascontiguousarray(x)) for x in (a1, A1.astype(np.float32), a2, A2.astype(np.float32), y)]
    loss = tr.step(*host, global_count=count)
    assert abs(loss.item() - float(o["loss"])) <= TOL * max(1.0, abs(float(o["loss"])))
    g = model.grad_dict()
    for k in sorted(o["grads"]):
        assert rel_err(g[k], o["grads"][k], floor=1e-6) <= TOL, k
    dev = [t.cuda() for t in host]
    loss2 = tr.step(*dev, global_count=count)
    assert abs(loss2.item() - loss.item()) <= 1e-6 * max(1.0, abs(loss.item()))


def test_config_D_inference_hidden256():
    """BASELINE config D at test scale: GGNN H256 T8 + GGNNReadout (R1) O=256 + HolE(hidden_dims=()) -> 1 logit,
    forward only (the fp32 kernels with four 64-channel chunks; N = 64 = the maximum the kernels take)."""
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(41)
    H, T, mb, N = 256, 8, 3, 64
    a1, A1 = synthetic.random_molecules(rng, mb, N)
    a2, A2 = synthetic.random_molecules(rng, mb, N)
    shapes = {"graph_conv/" + k: v for k, v in R.ggnn_shapes(H, H, T).items()}
    shapes.update({"mlp/" + k: v for k, v in R.hole_shapes(H, 1, ()).items()})
    params = R.init_params(shapes, rng, dtype=np.float64)
    tab = R.wrap_params(params)
    P = R.P(tab)
    omodel = R.GraphConvPredictorForPair(R.GGNN(P.sub("graph_conv"), H, H, T), None, R.HolE(P.sub("mlp"), 1, ()))
    ologits = omodel(a1, A1.astype(np.float64), a2, A2.astype(np.float64)).data
    model = gcnbmp.GraphConvPredictorForPair(gcnbmp.GGNN(H, H, T), None, gcnbmp.HolE(1, ()))
    model.load_params(params)
    with torch.no_grad():
        logits = model(a1, A1, a2, A2).cpu().numpy()
    assert rel_err(logits, ologits) <= TOL


def test_config_B_relgcn_64x4():
    """BASELINE config B at test scale: RelGCN 64->64 x4 layers, readout O=64, binary head, N = 64, fwd+bwd."""
    import gcnbmp
    case = cases.pair_case("B", seed=17)
    sp = dict(case["spec"], ch=[64, 64, 64, 64, 64], O=64, N1=64, N2=64, scale_adj=False)
    from gcnbmp import synthetic
    rng = np.random.default_rng(3)
    a1, A1 = synthetic.random_molecules(rng, 4, 64)
    a2, A2 = synthetic.random_molecules(rng, 4, 64)
    y = (rng.random((4, 1)) < 0.33).astype(np.int32)
    shapes = {"graph_conv/" + k: v for k, v in R.relgcn_shapes(64, sp["ch"]).items()}
    shapes.update({"mlp/" + k: v for k, v in R.hole_shapes(64, 1, ()).items()})
    big = dict(case, spec=sp, params=R.init_params(shapes, rng, dtype=np.float64),
               inputs=(a1, A1.astype(np.float64), a2, A2.astype(np.float64)), labels=y)
    o = cases.oracle_eval(big)
    p = product.product_eval(big)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    for k in sorted(o["grads"]):
        assert rel_err(p["grads"][k], o["grads"][k], floor=1e-7) <= TOL, k


def test_nfp_pair_forward_backward_matches_oracle():
    """train_binary.py --method nfp (the script's default encoder) + Nie co-attention + HolE: models/models/nfp.py on the
    (mb, N, N) adjacency with self connections, logits and every parameter gradient vs the fp64 oracle."""
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(12)
    mb, N, H, O, T, K = 5, 21, 16, 12, 3, 3
    ins = []
    for _ in range(2):
        atoms, adj4 = synthetic.random_molecules(rng, mb, N)
        adj = adj4.sum(axis=1) + np.eye(N, dtype=np.float32)[None] * (atoms > 0)[:, :, None]
        ins += [atoms, adj.astype(np.float64)]
    y = (rng.random((mb, K)) < 0.4).astype(np.int32)
    shapes = {"graph_conv/" + k: v for k, v in R.nfp_shapes(O, H, T).items()}
    shapes.update({"attn/" + k: v for k, v in R.coattn_shapes(H, O, 4).items()})
    shapes.update({"mlp/" + k: v for k, v in R.hole_shapes(O, K, ()).items()})
    params = R.init_params(shapes, rng, dtype=np.float64)
    tab = R.wrap_params(params)
    P = R.P(tab)
    omodel = R.GraphConvPredictorForPair(R.NFP(P.sub("graph_conv"), O, H, T), R.NieFineCoattention(P.sub("attn"), H, O, 4, activation="tanh"),
                                         R.HolE(P.sub("mlp"), K, hidden_dims=()))
    loss, logits, grads = R.loss_and_grads(omodel, tab, tuple(ins), y)
    model = gcnbmp.GraphConvPredictorForPair(gcnbmp.NFP(O, hidden_dim=H, n_layers=T),
                                             gcnbmp.NieFineCoattention(H, O, 4, activation=gcnbmp.functions.tanh), gcnbmp.HolE(K, hidden_dims=()))
    model.mlp.l_out.ensure(O)
    model.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
    model.cleargrads()
    f32 = [x.astype(np.float32) if x.dtype.kind == "f" else x for x in ins]
    plogits = model(*f32)
    gcnbmp.sigmoid_cross_entropy(plogits, y).backward()
    assert rel_err(plogits.detach().cpu().numpy(), logits) <= 1e-4
    g = model.grad_dict()
    for k, v in grads.items():
        if v is not None and np.abs(v).max() > 1e-9:
            assert rel_err(g[k], v) <= 1e-4, k


@pytest.mark.parametrize("mb,N1,N2,H,head", [(5, 64, 64, 32, 16), (3, 17, 40, 128, 8), (300, 9, 5, 16, 4)])
def test_bimpm_matches_oracle(mb, N1, N2, H, head):
    """models/coattention/bimpm.py (--attn bimpm) on csrc/bimpm.cu: outputs, atom gradients and the three weight gradients vs the
    fp64 oracle -- ragged N1 != N2, the script's head = out_dim = 16 shape, and more pairs than CTAs (grid-stride loop)."""
    import gcnbmp
    rng = np.random.default_rng(mb + N1)
    a1, a2 = rng.standard_normal((mb, N1, H)) * 0.6, rng.standard_normal((mb, N2, H)) * 0.6
    params = R.init_params(R.bimpm_shapes(H, head), rng, dtype=np.float64)
    w1, w2 = rng.standard_normal((mb, 3 * head)), rng.standard_normal((mb, 3 * head))
    tab = R.wrap_params(params)
    v1, v2 = F.param(a1), F.param(a2)
    o1, o2 = R.BiMPM(R.P(tab), H, 8, head)(v1, None, v2, None)
    F.add(F.sum_(F.mul(o1, F.const(w1))), F.sum_(F.mul(o2, F.const(w2)))).backward()
    net = gcnbmp.BiMPM(H, 8, head)
    net.load_params(params)
    t1 = torch.tensor(a1, dtype=torch.float32, device="cuda", requires_grad=True)
    t2 = torch.tensor(a2, dtype=torch.float32, device="cuda", requires_grad=True)
    p1, p2 = net(t1, None, t2, None)
    ((p1 * torch.tensor(w1, dtype=torch.float32, device="cuda")).sum() + (p2 * torch.tensor(w2, dtype=torch.float32, device="cuda")).sum()).backward()
    assert rel_err(p1.detach().cpu().numpy(), o1.data) <= TOL and rel_err(p2.detach().cpu().numpy(), o2.data) <= TOL
    assert rel_err(t1.grad.cpu().numpy(), v1.grad) <= TOL and rel_err(t2.grad.cpu().numpy(), v2.grad) <= TOL
    g = net.grad_dict()
    for k in params:
        assert rel_err(g[k], tab[k].grad) <= TOL, k


@pytest.mark.parametrize("H,T,tied,N", [(256, 3, True, 64), (192, 2, False, 21), (256, 2, True, 9)])
def test_fp32_training_at_hidden_above_128(H, T, tied, N):
    """Training at hidden 192 / 256 in BMP_MODE_F32 (ggnn_bwd_big_kernel): graph vectors, atom states and every parameter gradient of
    the GGNN encoder vs the fp64 oracle (the reference trains any --fp-hidden-dim, train_binary.py:297-429)."""
    import gcnbmp
    from gcnbmp import synthetic
    rng = np.random.default_rng(H + T)
    mb, O = 3, 40
    atoms, adj = synthetic.random_molecules(rng, mb, N)
    params = R.init_params(R.ggnn_mono_shapes(O, H, T, weight_tying=tied), rng, dtype=np.float64)
    tab = R.wrap_params(params)
    onet = R.GGNNMono(R.P(tab), O, H, T, weight_tying=tied)
    w_g, w_a = rng.standard_normal((mb, O)), rng.standard_normal((mb, N, H)) * 0.1
    og = onet(atoms, adj.astype(np.float64))
    oa = onet.get_atom_array()
    F.add(F.sum_(F.mul(og, F.const(w_g))), F.sum_(F.mul(oa, F.const(w_a)))).backward()
    net = gcnbmp.GGNNMono(O, H, T, weight_tying=tied)
    net.load_params(params)
    net.cleargrads()
    pg = net(atoms, adj)
    pa = net.get_atom_array()
    ((pg * torch.tensor(w_g, dtype=torch.float32, device="cuda")).sum() + (pa * torch.tensor(w_a, dtype=torch.float32, device="cuda")).sum()).backward()
    assert rel_err(pg.detach().cpu().numpy(), og.data) <= TOL and rel_err(pa.detach().cpu().numpy(), oa.data) <= TOL
    g = net.grad_dict()
    for k in params:
        if tab[k].grad is not None and np.abs(tab[k].grad).max() > 1e-12:
            assert rel_err(g[k], tab[k].grad) <= TOL, k


def test_config_d_model_trains_at_hidden_256_in_fp32():
    """BASELINE config D's model (modular GGNN hidden 256 + R1 read-out + HolE, no co-attention) as a TRAINING step in fp32 mode:
    logits and every parameter gradient vs the fp64 oracle."""
    case = cases.pair_case("MU", seed=3)
    sp = dict(case["spec"], H=256, O=64, T=2, tied=True, hole_hidden=())
    rng = np.random.default_rng(21)
    shapes = {"graph_conv/" + k: v for k, v in R.ggnn_shapes(64, 256, 2, weight_tying=True).items()}
    shapes.update({"mlp/" + k: v for k, v in R.hole_shapes(64, sp["K"], ()).items()})
    big = dict(case, spec=sp, params=R.init_params(shapes, rng, dtype=np.float64))
    o = cases.oracle_eval(big)
    p = product.product_eval(big)
    assert rel_err(p["logits"], o["logits"]) <= TOL
    for k, v in o["grads"].items():
        if v is not None and np.abs(v).max() > 1e-12:
            assert rel_err(p["grads"][k], v) <= TOL, k


@pytest.mark.parametrize("H,T,tied,N,mb,weighted", [(128, 3, True, 64, 5, False), (128, 2, False, 37, 9, True), (64, 3, True, 50, 7, False),
                                                      (256, 2, True, 30, 6, False)])
def test_fp32_mode_on_tensor_cores_matches_oracle_and_the_ffma_kernels(H, T, tied, N, mb, weighted):
    """BMP_MODE_F32 with the encoder's contractions on tcgen05 (csrc/ggnn_x3.cu: bf16 hi/lo split, three UMMAs per product) against
    the fp64 oracle at the mode's 1e-4 bound, against the FFMA kernels of csrc/ggnn.cu (same stash, same results up to rounding), and
    its stash-free inference path against its training path.  Row counts are not multiples of the 128-row tile (tail tiles)."""
    import gcnbmp
    from gcnbmp import functional as Fn, synthetic
    rng = np.random.default_rng(H + T + N)
    O = 40
    atoms, adj = synthetic.random_molecules(rng, mb, N)
    if weighted:        # general fp32 edge weights (not symmetric): the adjacency bit masks only say WHERE the entries are
        adj = (adj * rng.uniform(0.25, 1.5, size=adj.shape)).astype(np.float32)
    assert mb * N >= 128 and (mb * N) % 128 != 0
    params = R.init_params(R.ggnn_mono_shapes(O, H, T, weight_tying=tied), rng, dtype=np.float64)
    tab = R.wrap_params(params)
    onet = R.GGNNMono(R.P(tab), O, H, T, weight_tying=tied)
    w_g, w_a = rng.standard_normal((mb, O)), rng.standard_normal((mb, N, H)) * 0.1
    og = onet(atoms, adj.astype(np.float64))
    oa = onet.get_atom_array()
    F.add(F.sum_(F.mul(og, F.const(w_g))), F.sum_(F.mul(oa, F.const(w_a)))).backward()
    got = {}
    try:
        for tc in (True, False):
            Fn.F32_TENSOR_CORES = tc
            net = gcnbmp.GGNNMono(O, H, T, weight_tying=tied)
            net.load_params(params)
            net.cleargrads()
            launches0 = gcnbmp._capi.lib.bmp_launch_count()
            pg = net(atoms, adj)
            pa = net.get_atom_array()
            ((pg * torch.tensor(w_g, dtype=torch.float32, device="cuda")).sum() + (pa * torch.tensor(w_a, dtype=torch.float32, device="cuda")).sum()).backward()
            got[tc] = dict(g=pg.detach().cpu().numpy(), a=pa.detach().cpu().numpy(), grads=net.grad_dict(),
                           launches=gcnbmp._capi.lib.bmp_launch_count() - launches0)
            assert rel_err(got[tc]["g"], og.data) <= TOL and rel_err(got[tc]["a"], oa.data) <= TOL
            for k in params:
                if tab[k].grad is not None and np.abs(tab[k].grad).max() > 1e-12:
                    assert rel_err(got[tc]["grads"][k], tab[k].grad) <= TOL, (tc, k)
            with torch.no_grad():
                ig = net(atoms, adj)
                ia = net.get_atom_array()
            assert rel_err(ig.cpu().numpy(), got[tc]["g"]) <= 1e-6 and rel_err(ia.cpu().numpy(), got[tc]["a"]) <= 1e-6
    finally:
        Fn.F32_TENSOR_CORES = True
    # the two paths are different kernels (a launch per GEMM vs one fused launch) and agree far inside the bound
    assert got[True]["launches"] > got[False]["launches"] + 4 * T
    assert rel_err(got[True]["a"], got[False]["a"]) <= 5e-5
    for k in params:
        if np.abs(got[False]["grads"][k]).max() > 1e-12:
            assert rel_err(got[True]["grads"][k], got[False]["grads"][k]) <= 5e-5, k


@pytest.mark.parametrize("name,mode", [("C", "f32"), ("CB", "f32"), ("CB", "bf16"), ("M2", "f32")])
def test_one_call_pair_step_matches_oracle_and_the_composed_path(name, mode):
    """bmp_pair_forward_backward (csrc/pair.cu): the whole pair step -- both encoder passes, co-attention, HolE, sigmoid-CE and the
    backward chain into caller-owned gradient buffers -- as ONE C call, against the fp64 oracle (fp32 mode: 1e-4) and against the
    composed autograd path through the same kernels (both modes)."""
    import gcnbmp
    case = cases.pair_case(name, seed=11)
    a1, A1, a2, A2 = case["inputs"]
    A1, A2 = A1.astype(np.float32), A2.astype(np.float32)
    y = case["labels"]
    o = cases.oracle_eval(case)
    ref = product.product_eval(case, mode=gcnbmp.MODE_BF16 if mode == "bf16" else gcnbmp.MODE_F32)
    model = product.product_model(case["spec"], case["params"])
    if mode == "bf16":
        model.graph_conv.mode = model.attn.mode = gcnbmp.MODE_BF16
    assert gcnbmp.fused.supported(model)
    model.cleargrads()
    n0 = gcnbmp.launch_count()
    loss, logits = gcnbmp.fused.pair_forward_backward(model, a1, A1, a2, A2, y)
    assert gcnbmp.launch_count() > n0
    grads = model.grad_dict()
    tol_ref = 1e-5 if mode == "f32" else 2e-3          # same kernels; bf16: atomics order + tape differences
    assert rel_err(logits.cpu().numpy(), ref["logits"]) <= tol_ref
    assert abs(float(loss) - ref["loss"]) <= tol_ref * max(1.0, abs(ref["loss"]))
    for k in sorted(ref["grads"]):
        assert rel_err(grads[k], ref["grads"][k]) <= max(tol_ref, 2e-5), (k, rel_err(grads[k], ref["grads"][k]))
    if mode == "f32":
        assert rel_err(logits.cpu().numpy(), o["logits"]) <= TOL
        assert abs(float(loss) - float(o["loss"])) <= TOL * max(1.0, abs(float(o["loss"])))
        for k in sorted(o["grads"]):
            assert rel_err(grads[k], o["grads"][k]) <= TOL, (k, rel_err(grads[k], o["grads"][k]))


def test_one_call_pair_step_refuses_a_short_workspace_and_unsupported_models():
    import ctypes as C
    import gcnbmp
    from gcnbmp import _capi as K
    a = K.Pair()
    a.mb, a.n1, a.n2, a.hidden, a.out_dim, a.head, a.n_classes, a.n_steps, a.n_atom_types, a.mode = 4, 20, 20, 32, 16, 8, 3, 2, 117, 0
    buf = torch.zeros(4096, device="cuda")
    for f in ("atoms_1", "atoms_2", "adj_1", "adj_2", "labels", "embed_W", "out_W", "logits", "loss", "workspace"):
        setattr(a, f, C.c_void_p(buf.data_ptr()))
    a.count, a.workspace_bytes = 12.0, 1024
    rc = K.lib.bmp_pair_forward_backward(C.byref(a), None)
    assert rc == -1 and b"workspace" in K.lib.bmp_last_error()
    case = cases.pair_case("MU", seed=1)            # modular GGNN + HolE without attention: outside the one-call composition
    model = product.product_model(case["spec"], case["params"])
    assert not gcnbmp.fused.supported(model)
    with pytest.raises(ValueError):
        gcnbmp.fused.pair_forward_backward(model, *[x for x in case["inputs"]], case["labels"])


@pytest.mark.parametrize("H,O,variant,use_h0,use_mask,nobias,act", [(128, 128, "R2", True, False, False, "identity"),
                                                                    (64, 256, "R1", True, True, False, "tanh"),
                                                                    (256, 64, "R1", False, False, True, "relu"),
                                                                    (128, 128, "R1", True, False, False, "identity")])
def test_fp32_readout_forward_on_tensor_cores_matches_oracle_and_the_ffma_kernel(H, O, variant, use_h0, use_mask, nobias, act):
    """BMP_MODE_F32 read-out forward with both linears as split-bf16 row GEMMs (csrc/ggnn_x3.cu) vs the fp64 oracle formulas
    (ggnn_readout.py:42-58 / ggnn_att.py:338-346) and vs the FFMA kernel; the backward (unchanged kernel) through it."""
    import gcnbmp
    from gcnbmp import functional as Fn, _capi as K
    rng = np.random.default_rng(H + O)
    mb, N = 7, 41                                     # 287 rows: two full tiles and a tail
    h, h0 = rng.standard_normal((mb, N, H)) * 0.5, rng.standard_normal((mb, N, H)) * 0.5
    mask = (rng.random((mb, N)) < 0.7).astype(np.float64) if use_mask else None
    kin_i = 2 * H if use_h0 else H
    kin_j = H if variant == "R2" else kin_i
    Wi, Wj = rng.standard_normal((O, kin_i)) / np.sqrt(kin_i), rng.standard_normal((O, kin_j)) / np.sqrt(kin_j)
    bi, bj = (None, None) if nobias else (rng.standard_normal(O) * 0.1, rng.standard_normal(O) * 0.1)
    A = {"identity": lambda x: x, "tanh": np.tanh, "relu": lambda x: np.maximum(x, 0)}[act]
    h1 = np.concatenate([h, h0], axis=2) if use_h0 else h
    gi = 1.0 / (1.0 + np.exp(-(h1 @ Wi.T + (0 if bi is None else bi))))
    gj = A((h if variant == "R2" else h1) @ Wj.T + (0 if bj is None else bj))
    g = gi * gj
    if mask is not None:
        g = g * mask[:, :, None]
    ref = A(g.sum(axis=1)) if variant == "R1" else g.sum(axis=1)
    t = lambda x: None if x is None else torch.tensor(x, dtype=torch.float32, device="cuda")
    code = dict(identity=0, tanh=1, relu=2)[act]
    args = (t(h), t(h0) if use_h0 else None, t(mask), K.READOUT_R1 if variant == "R1" else K.READOUT_R2, code,
            code if variant == "R1" else 0, t(Wi), t(bi), t(Wj), t(bj), K.MODE_F32)
    out = {}
    try:
        for tc in (True, False):
            Fn.F32_TENSOR_CORES = tc
            n0 = gcnbmp.launch_count()
            with torch.no_grad():
                out[tc] = Fn.readout(*args).cpu().numpy()
            out[(tc, "n")] = gcnbmp.launch_count() - n0
            assert rel_err(out[tc], ref) <= TOL, (tc, rel_err(out[tc], ref))
    finally:
        Fn.F32_TENSOR_CORES = True
    assert out[(True, "n")] > out[(False, "n")]            # pack + GEMM + reduce vs one fused FFMA launch
    assert rel_err(out[True], out[False]) <= 2e-5
    # training through the tensor-core forward: the backward kernel sees the same g
    hh = t(h).requires_grad_()
    gg = Fn.readout(hh, *args[1:])
    gg.sum().backward()
    assert torch.isfinite(hh.grad).all() and float(hh.grad.abs().max()) > 0


@pytest.mark.parametrize("storage", ["u8", "bits"])
def test_one_call_pair_step_takes_byte_and_bit_packed_adjacency_in_bf16_mode(storage):
    import gcnbmp
    case = cases.pair_case("CB", seed=3)
    a1, A1, a2, A2 = case["inputs"]
    y = case["labels"]
    conv = (lambda A: A.astype(np.uint8)) if storage == "u8" else (lambda A: gcnbmp.pack_adjacency(A.astype(np.float32)))
    res = []
    for adjs in ((A1.astype(np.float32), A2.astype(np.float32)), (conv(A1), conv(A2))):
        model = product.product_model(case["spec"], case["params"])
        model.graph_conv.mode = model.attn.mode = gcnbmp.MODE_BF16
        model.cleargrads()
        loss, logits = gcnbmp.fused.pair_forward_backward(model, a1, adjs[0], a2, adjs[1], y)
        res.append((float(loss), logits.cpu().numpy(), model.grad_dict()))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])          # same staged bf16 tiles: bit-identical forward
    for k in res[0][2]:
        assert rel_err(res[1][2][k], res[0][2][k]) <= 1e-3, k                          # gradients: atomics order only
