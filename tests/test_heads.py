"""The remaining link-prediction heads (MLP / SymMLP / NTN / DistMult, models/mlp.py) and the optimizer hooks
(train_binary.py:537-543).  CPU: the oracle's op-sequence restatement against an independent closed-form Torch-autograd
twin in fp64.  GPU: the CUDA heads / hook kernel against the oracle (fp32 parity bound 1e-4)."""
import numpy as np
import pytest
import torch

from oracle import minichainer as F
from oracle import reference_path as R

KINDS = [("mlp", (32, 16)), ("symmlp", (32,)), ("ntn", (16,)), ("distmult", ()), ("ntn", ())]


def _oracle_head(kind, tab, D, K, hidden):
    p = R.P(tab)
    if kind == "mlp":
        net = R.MLP(p, K, hidden)
        return lambda l, r: net(F.concat((l, r), axis=-1))
    if kind == "symmlp":
        return R.SymMLP(p, K, hidden)
    if kind == "ntn":
        return R.NTN(p, D, D, K, 8, hidden)
    return R.DistMult(p, D, D, K, 8, hidden)


def _twin_head(kind, P, hidden):
    """closed forms, written independently of the oracle's op sequence"""
    names = "layers" if kind in ("mlp", "symmlp") else "mlp_layers"

    def stack(h):
        for i in range(len(hidden)):
            h = torch.relu(h @ P["%s/%d/W" % (names, i)].T + P["%s/%d/b" % (names, i)])
        return h @ P["l_out/W"].T + P["l_out/b"]

    def fn(l, r):
        if kind == "mlp":
            return stack(torch.cat([l, r], dim=1))
        if kind == "symmlp":
            return stack(torch.cat([l + r, l * r], dim=1))
        if kind == "ntn":
            h = (torch.einsum("bi,ijk,bj->bk", l, P["ntn_layer/W"], r) + l @ P["ntn_layer/V1"] + r @ P["ntn_layer/V2"] + P["ntn_layer/b"])
            return stack(h)
        # models/mlp.py:188 builds the diagonal slices as a float32 constant
        return stack(torch.einsum("bi,ki,bi->bk", l, P["dm_layer/W"].detach().float().double(), r))
    return fn


def _case(kind, hidden, seed=0, mb=7, D=24, K=3):
    rng = np.random.default_rng(seed)
    params = R.init_params(R.head_shapes(kind, D, K, hidden), rng, dtype=np.float64)
    l, r = rng.standard_normal((mb, D)), rng.standard_normal((mb, D))
    w = rng.standard_normal((mb, K))
    return params, l, r, w


def _oracle_eval(kind, hidden, params, l, r, w, D, K):
    tab = R.wrap_params(params)
    lv, rv = F.param(l), F.param(r)
    out = _oracle_head(kind, tab, D, K, hidden)(lv, rv)
    F.sum_(F.mul(out, F.const(w))).backward()
    return out.data, lv.grad, rv.grad, {k: v.grad for k, v in tab.items()}


@pytest.mark.parametrize("kind,hidden", KINDS)
def test_oracle_heads_match_closed_form_twin(kind, hidden):
    D, K = 24, 3
    params, l, r, w = _case(kind, hidden)
    out, gl, gr, gp = _oracle_eval(kind, hidden, params, l, r, w, D, K)
    P = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    lt, rt = torch.tensor(l, requires_grad=True), torch.tensor(r, requires_grad=True)
    tout = _twin_head(kind, P, hidden)(lt, rt)
    (tout * torch.tensor(w)).sum().backward()
    np.testing.assert_allclose(out, tout.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gl, lt.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gr, rt.grad.numpy(), rtol=1e-9, atol=1e-12)
    for k in params:
        if k == "dm_layer/W":
            assert gp[k] is None and P[k].grad is None       # models/mlp.py:186-192: built from W.data, no gradient
        else:
            np.testing.assert_allclose(gp[k], P[k].grad.numpy(), rtol=1e-9, atol=1e-12, err_msg=k)


def test_symmlp_is_symmetric_and_distmult_too():
    for kind, hidden in (("symmlp", (8,)), ("distmult", (4,))):
        params, l, r, w = _case(kind, hidden, seed=3)
        a = _oracle_eval(kind, hidden, params, l, r, w, 24, 3)[0]
        b = _oracle_eval(kind, hidden, params, r, l, w, 24, 3)[0]
        np.testing.assert_allclose(a, b, rtol=1e-12)


def test_oracle_hooks_closed_forms():
    rng = np.random.default_rng(1)
    p = {"a": rng.standard_normal((5, 3)), "b": rng.standard_normal((7,))}
    g = {"a": rng.standard_normal((5, 3)) * 3, "b": rng.standard_normal((7,)) * 3}
    out = R.apply_hooks(g, p, max_norm=1.5)
    norm = np.sqrt(sum((v ** 2).sum() for v in out.values()))
    assert abs(norm - 1.5) < 1e-12                                     # clipped onto the ball
    big = R.apply_hooks(g, p, max_norm=1e6)
    np.testing.assert_array_equal(big["a"], g["a"])                    # rate >= 1: untouched
    out = R.apply_hooks(g, p, l2_rate=0.1, l1_rate=0.01)
    np.testing.assert_allclose(out["b"], g["b"] + 0.1 * p["b"] + 0.01 * np.sign(p["b"]))
    out = R.apply_hooks(g, p, max_norm=1.5, l2_rate=0.1)              # clipping first, decay on the clipped gradient
    rate = 1.5 / np.sqrt(sum((v ** 2).sum() for v in g.values()))
    np.testing.assert_allclose(out["a"], g["a"] * rate + 0.1 * p["a"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kind,hidden", KINDS)
@pytest.mark.parametrize("mb,D,K", [(7, 24, 3), (300, 128, 86), (5, 33, 1)])
def test_cuda_heads_match_oracle(kind, hidden, mb, D, K):
    import gcnbmp
    from product import rel_err
    rng = np.random.default_rng(mb + D)
    params = R.init_params(R.head_shapes(kind, D, K, hidden), rng, dtype=np.float64)
    l, r, w = rng.standard_normal((mb, D)), rng.standard_normal((mb, D)), rng.standard_normal((mb, K))
    out, gl, gr, gp = _oracle_eval(kind, hidden, params, l, r, w, D, K)
    link = {"mlp": lambda: gcnbmp.MLP(K, hidden), "symmlp": lambda: gcnbmp.SymMLP(K, hidden),
            "ntn": lambda: gcnbmp.NTN(D, D, K, 8, hidden), "distmult": lambda: gcnbmp.DistMult(D, D, K, 8, hidden)}[kind]()
    link.load_params(params)
    lt = torch.tensor(l, dtype=torch.float32, device="cuda", requires_grad=True)
    rt = torch.tensor(r, dtype=torch.float32, device="cuda", requires_grad=True)
    if kind == "mlp":
        pout = link(gcnbmp.functional.PairFeatures.apply(lt, rt, gcnbmp._capi.PAIR_CONCAT))
    else:
        pout = link(lt, rt)
    (pout * torch.tensor(w, dtype=torch.float32, device="cuda")).sum().backward()
    assert rel_err(pout.detach().cpu().numpy(), out) <= 1e-4
    assert rel_err(lt.grad.cpu().numpy(), gl) <= 1e-4 and rel_err(rt.grad.cpu().numpy(), gr) <= 1e-4
    g = link.grad_dict()
    for k in params:
        if k == "dm_layer/W":
            assert g.get(k) is None or not np.any(g[k])
        else:
            assert rel_err(g[k], gp[k]) <= 1e-4, k


@pytest.mark.gpu
def test_pair_predictor_with_ntn_head_matches_oracle():
    """GGNN encoder + co-attention + NTN head through the pair predictor (train_binary.py:84-118), fp32 parity."""
    import cases
    import gcnbmp
    from product import rel_err
    case = cases.pair_case("C", seed=4)
    sp = case["spec"]
    H = O = sp["O"]
    rng = np.random.default_rng(9)
    params = {k: v for k, v in case["params"].items() if not k.startswith("mlp/")}
    params.update(R.init_params({"mlp/" + k: v for k, v in R.head_shapes("ntn", O, sp["K"], (16,)).items()}, rng, dtype=np.float64))
    tab = R.wrap_params(params)
    enc = R.GGNNMono(R.P(tab).sub("graph_conv"), O, sp["H"], sp["T"])
    attn = R.NieFineCoattention(R.P(tab).sub("attn"), sp["H"], O, sp["head"], activation="tanh")
    opred = R.GraphConvPredictorForPair(enc, attn, R.NTN(R.P(tab).sub("mlp"), O, O, sp["K"], 8, (16,)))
    loss, logits, grads = R.loss_and_grads(opred, tab, [np.asarray(x, np.float64) if x.dtype.kind == "f" else x for x in case["inputs"]], case["labels"])
    f = gcnbmp.functions
    model = gcnbmp.GraphConvPredictorForPair(gcnbmp.GGNNMono(O, sp["H"], sp["T"]),
                                             gcnbmp.NieFineCoattention(sp["H"], O, sp["head"], activation=f.tanh),
                                             gcnbmp.NTN(O, O, sp["K"], 8, (16,)))
    a1, A1, a2, A2 = case["inputs"]
    plogits = model(a1, A1.astype(np.float32), a2, A2.astype(np.float32))     # materialises lazy layers
    model.load_params(params)
    model.cleargrads()
    plogits = model(a1, A1.astype(np.float32), a2, A2.astype(np.float32))
    ploss = gcnbmp.sigmoid_cross_entropy(plogits, case["labels"])
    ploss.backward()
    assert rel_err(plogits.detach().cpu().numpy(), logits) <= 1e-4
    g = model.grad_dict()
    for k in grads:
        if grads[k] is not None and np.abs(grads[k]).max() > 1e-9:
            assert rel_err(g[k], grads[k]) <= 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("max_norm,l2,l1", [(1.5, 0.0, 0.0), (1e6, 0.0, 0.0), (0.0, 0.1, 0.01), (2.0, 0.05, 0.02)])
def test_gradient_hooks_kernel_matches_chainer_semantics(max_norm, l2, l1):
    import ctypes as C
    import gcnbmp
    from gcnbmp import train
    rng = np.random.default_rng(2)
    n = 100003
    p, g = rng.standard_normal(n).astype(np.float32), (rng.standard_normal(n) * 0.05).astype(np.float32)
    p[:10] = 0.0                                                         # sign(0) = 0
    ref = R.apply_hooks({"x": g}, {"x": p.astype(np.float64)}, max_norm, l2, l1)["x"]
    pt, gt = torch.tensor(p).cuda(), torch.tensor(g).cuda()
    hooks = train.GradientHooks(pt, gt, max_norm, l2, l1)
    assert hooks.active
    hooks.apply()
    np.testing.assert_allclose(gt.cpu().numpy(), ref, rtol=2e-5, atol=1e-7)


@pytest.mark.gpu
def test_trainer_applies_hooks_before_adam_and_exponential_shift():
    import cases
    import gcnbmp
    import product
    from gcnbmp import train
    case = cases.pair_case("A", seed=3)
    a1, A1, a2, A2 = case["inputs"]
    y = case["labels"]

    def run(**kw):
        model = product.product_model(case["spec"], case["params"])
        tr = train.PairTrainer(model, chunk=64, alpha=1e-2, **kw)
        tr.step(a1, A1.astype(np.float32), a2, A2.astype(np.float32), y)
        return tr, tr.flat.detach().cpu().numpy().copy(), tr.gflat.detach().cpu().numpy().copy()

    _, p0, g0 = run()
    tr, p1, g1 = run(max_norm=1e-3, l2_rate=0.1)
    assert np.linalg.norm(g0) > 1e-3
    assert not np.allclose(p0, p1)
    # the flat gradient after the step is the hooked one: clipped to the ball, then decayed
    start = product.product_model(case["spec"], case["params"]).flatten_parameters()[0].detach().cpu().numpy()
    np.testing.assert_allclose(g1, g0 * (1e-3 / np.linalg.norm(g0)) + 0.1 * start, rtol=1e-4, atol=1e-7)
    shift = train.ExponentialShift(tr.opt, 0.5, epochs=[2, 4])
    assert shift.maybe(1) is None and abs(shift.maybe(2) - 5e-3) < 1e-12 and shift.maybe(2) is None
    assert abs(shift.maybe(4) - 2.5e-3) < 1e-12 and abs(tr.opt.hp[0] - 2.5e-3) < 1e-12


@pytest.mark.gpu
def test_batch_evaluator_scores_a_split_once_and_matches_sklearn():
    """f-3: one forward pass over the split, metrics on the device == scikit-learn on the same sigmoid(logits)."""
    sk = pytest.importorskip("sklearn.metrics")
    import cases
    import gcnbmp
    import product
    from gcnbmp import synthetic
    from gcnbmp.evaluate import BatchEvaluator
    case = cases.pair_case("A", seed=5)
    model = product.product_model(case["spec"], case["params"])
    rng = np.random.default_rng(0)
    n = 300
    a1, A1 = synthetic.random_molecules(rng, n, 50)
    a2, A2 = synthetic.random_molecules(rng, n, 50)
    y = (rng.random((n, 1)) < 0.4).astype(np.int32)
    y[0], y[1] = 1, 0
    ev = BatchEvaluator(model, name="val", batch=128)
    gcnbmp.reset_launch_count()
    obs = ev.evaluate(a1, A1, a2, A2, y)
    with torch.no_grad():
        logits = torch.cat([model(a1[s:s + 128], A1[s:s + 128], a2[s:s + 128], A2[s:s + 128]) for s in range(0, n, 128)])
    p = torch.sigmoid(logits.double()).cpu().numpy()
    pr, rc, _ = sk.precision_recall_curve(y[:, 0], p[:, 0], pos_label=1)
    ref = {"val/roc_auc": sk.roc_auc_score(y, p), "val/prc_auc": sk.auc(rc, pr), "val/accuracy": sk.accuracy_score(y, np.round(p)),
           "val/f1": sk.f1_score(y, np.round(p), average="macro", zero_division=0)}
    for k, v in ref.items():
        assert abs(obs[k] - v) <= 1e-6, (k, obs[k], v)
    # table form: the same pairs as index pairs give the same numbers
    tab_a, tab_A = torch.tensor(np.concatenate([a1, a2])).cuda(), torch.tensor(np.concatenate([A1, A2])).cuda()
    obs2 = ev.evaluate_indexed(tab_a, tab_A, np.arange(n), n + np.arange(n), y)
    for k in ref:
        assert abs(obs2[k] - obs[k]) <= 1e-5, k


@pytest.mark.gpu
def test_trainer_and_evaluator_share_one_flat_buffer():
    """flatten_parameters is idempotent: an evaluator built on a model that a trainer already owns must not re-home the
    parameters (the optimiser would keep updating buffers the model no longer reads)."""
    import cases
    import product
    from gcnbmp import train
    from gcnbmp.evaluate import BatchEvaluator
    case = cases.pair_case("A", seed=3)
    a1, A1, a2, A2 = case["inputs"]
    A1, A2, y = A1.astype(np.float32), A2.astype(np.float32), case["labels"]
    model = product.product_model(case["spec"], case["params"])
    tr = train.PairTrainer(model, chunk=64, alpha=1e-2)
    tr.step(a1, A1, a2, A2, y)
    p1 = tr.flat.detach().clone()
    ev = BatchEvaluator(model, name="val", batch=64, which=("accuracy",))
    assert ev.trainer.flat.data_ptr() == tr.flat.data_ptr() and ev.trainer.gflat.data_ptr() == tr.gflat.data_ptr()
    before = ev.evaluate(a1, A1, a2, A2, y)
    tr.step(a1, A1, a2, A2, y)
    assert not torch.equal(tr.flat, p1)                      # the second step still moved the parameters ...
    named = dict(model.namedparams())
    k = next(iter(named))
    assert named[k].data_ptr() >= tr.flat.data_ptr() and named[k].data_ptr() < tr.flat.data_ptr() + 4 * tr.flat.numel()
    with torch.no_grad():                                     # ... and the model computes with the updated ones
        l_model = model(a1, A1, a2, A2)
        l_pred = tr.predict(a1, A1, a2, A2)
    assert torch.equal(l_model, l_pred)
    assert isinstance(before["val/accuracy"], float)


@pytest.mark.gpu
def test_hooks_and_adam_skip_parameters_without_gradient():
    """BilinearDiag.W (DistMult) gets no gradient in the reference, so chainer's hooks / weight decay / Adam never touch it."""
    import gcnbmp
    from gcnbmp import synthetic, train
    rng = np.random.default_rng(1)
    a1, A1 = synthetic.random_molecules(rng, 6, 20)
    a2, A2 = synthetic.random_molecules(rng, 6, 20)
    y = (rng.random((6, 1)) < 0.5).astype(np.int32)
    model = gcnbmp.GraphConvPredictorForPair(gcnbmp.GGNNMono(16, 16, 2), None, gcnbmp.DistMult(16, 16, 1, 4, (6,)))
    with torch.no_grad():
        model(a1, A1, a2, A2)                                 # materialise the lazily-shaped layers
    tr = train.PairTrainer(model, chunk=64, alpha=1e-2, l2_rate=0.1, l1_rate=0.01, weight_decay_rate=0.1)
    w0 = model.mlp.dm_layer.W.detach().clone()
    e0 = model.graph_conv.embed.W.detach().clone()
    tr.step(a1, A1, a2, A2, y)
    assert torch.equal(model.mlp.dm_layer.W.detach(), w0)     # frozen, not decayed
    assert not torch.equal(model.graph_conv.embed.W.detach(), e0)


@pytest.mark.gpu
def test_default_global_count_ignores_missing_labels():
    """PairTrainer.step without `global_count` divides by the number of non-ignored labels, as F.sigmoid_cross_entropy does."""
    import cases
    import gcnbmp
    import product
    from gcnbmp import train
    case = cases.pair_case("C", seed=2)                       # K = 86 with one ignored entry
    a1, A1, a2, A2 = case["inputs"]
    A1, A2, y = A1.astype(np.float32), A2.astype(np.float32), case["labels"]
    assert (y == -1).sum() == 1
    model = product.product_model(case["spec"], case["params"])
    tr = train.PairTrainer(model, chunk=64, optimizer=False)
    loss = float(tr.step(a1, A1, a2, A2, y).item())
    with torch.no_grad():
        ref = float(gcnbmp.sigmoid_cross_entropy(model(a1, A1, a2, A2), y).item())
    assert abs(loss - ref) <= 1e-6 * max(abs(ref), 1.0)


@pytest.mark.gpu
def test_batch_evaluator_ignore_labels_with_several_columns():
    sk = pytest.importorskip("sklearn.metrics")
    import cases
    import product
    from gcnbmp.evaluate import BatchEvaluator
    case = cases.pair_case("C", seed=9)
    model = product.product_model(case["spec"], case["params"])
    a1, A1, a2, A2 = case["inputs"]
    rng = np.random.default_rng(3)
    reps = 10
    a1, A1, a2, A2 = (np.concatenate([x] * reps) for x in (a1, A1.astype(np.float32), a2, A2.astype(np.float32)))
    n, k = a1.shape[0], case["spec"]["K"]
    y = (rng.random((n, k)) < 0.4).astype(np.int32)
    y[0], y[1] = 1, 0
    y[rng.random((n, k)) < 0.15] = -1
    y[:2] = np.where(y[:2] < 0, 0, y[:2])
    y[0], y[1] = 1, 0
    obs = BatchEvaluator(model, name="val", batch=16, which=("roc_auc", "accuracy")).evaluate(a1, A1, a2, A2, y)
    with torch.no_grad():
        p = torch.sigmoid(model(a1, A1, a2, A2).double()).cpu().numpy()
    roc = np.mean([sk.roc_auc_score(y[y[:, c] >= 0, c], p[y[:, c] >= 0, c]) for c in range(k)])
    acc = np.mean([sk.accuracy_score(y[y[:, c] >= 0, c], np.round(p[y[:, c] >= 0, c])) for c in range(k)])
    assert abs(obs["val/roc_auc"] - roc) <= 1e-6 and abs(obs["val/accuracy"] - acc) <= 1e-6
