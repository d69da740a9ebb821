"""Generate tests/golden/ref_*.npz by running the REFERENCE'S OWN link files (/root/reference/models/..., unmodified, loaded
by path) over oracle/chainer_shim (stand-ins for the chainer / chainer_chemistry primitives on the oracle's NumPy tape).
Each fixture holds parameters (Chainer paths), inputs, a random cotangent w, the outputs and the gradients of sum(w * out)
with respect to every parameter and float input, all in float64.  The script also asserts that oracle/reference_path.py
reproduces every number to 1e-10 before writing.  Run in the build container only (the reference does not travel):
    python tests/golden/make_reference_golden.py
tests/test_reference_golden.py checks the oracle against the committed fixtures everywhere else."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GCNBMP_REFERENCE", "/root/reference")
for p in (os.path.join(ROOT, "gcn-bmp_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "chainer_shim"), ROOT):
    sys.path.insert(0, p)

from oracle import minichainer as M            # noqa: E402
from oracle import reference_path as R         # noqa: E402
import chainer                                  # noqa: E402  (the shim)
from chainer import functions as CF             # noqa: E402


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# the reference's packages use Python-2 implicit relative imports in their __init__ files; the hot-path FILES are Python-3
# clean, so they are loaded by path and the package names they import from are provided as empty modules
for pkg in ("models", "models.update", "models.readout", "update", "readout"):
    sys.modules[pkg] = types.ModuleType(pkg)
ref_ggnn_update = load("models.update.ggnn_update", "models/update/ggnn_update.py")
ref_relgcn_update = load("models.update.relgcn_update", "models/update/relgcn_update.py")
ref_ggnn_readout = load("models.readout.ggnn_readout", "models/readout/ggnn_readout.py")
for pkg in ("models.update", "update"):
    sys.modules[pkg].GGNNUpdate, sys.modules[pkg].RelGCNUpdate = ref_ggnn_update.GGNNUpdate, ref_relgcn_update.RelGCNUpdate
for pkg in ("models.readout", "readout"):
    sys.modules[pkg].GGNNReadout = ref_ggnn_readout.GGNNReadout
ref_ggnn = load("ref_models_ggnn", "models/models/ggnn.py")
ref_relgcn = load("ref_relgcn", "models/relgcn.py")
ref_nie = load("ref_nie", "models/coattention/nie_coattention.py")
ref_vqa = load("ref_vqa", "models/coattention/vqa_parallel_coattention.py")
ref_pool = load("ref_pool", "models/coattention/PoolingFineCoattention.py")
ref_alter = load("ref_alter", "models/coattention/alternating_coattention.py")
ref_para = load("ref_para", "models/coattention/parallel_coattention.py")
ref_global = load("ref_global", "models/coattention/global_coattention.py")
ref_neural = load("ref_neural", "models/coattention/neural_coattention.py")
ref_gin = load("ref_gin", "models/gin.py")
ref_nfp = load("ref_nfp", "models/models/nfp.py")
ref_bimpm = load("ref_bimpm", "models/coattention/bimpm.py")
ref_hole = load("ref_hole", "models/link_prediction/hole.py")
ref_mlp = load("ref_mlp", "models/mlp.py")
ref_mono = {"ggnn_py": load("ref_ggnn_py", "models/ggnn.py"), "ggnn_att": load("ref_ggnn_att", "models/ggnn_att.py"),
            "ggnn_dev": load("ref_ggnn_dev", "models/ggnn_dev.py")}


def load_class(rel, cls_name):
    """The reference's training scripts cannot be imported (argparse / RDKit / trainer set-up at module level); the source text of
    ONE class is taken from the file, unmodified, and executed against the shim."""
    import ast
    src = open(os.path.join(REF, rel)).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    ns = {"chainer": chainer, "cuda": chainer.cuda, "F": CF, "L": chainer.links}
    exec(compile(ast.Module(body=[node], type_ignores=[]), os.path.join(REF, rel), "exec"), ns)
    return ns[cls_name]


def load_params(link, table):
    """Chainer-path table -> the shim link's parameters (shapes checked; every table entry must be consumed)."""
    seen = set()
    for path, p in list(link.namedparams()):
        key = path.lstrip("/")
        assert key in table, "reference link has a parameter the table lacks: %s" % key
        owner, parts = link, key.split("/")
        for c in parts[:-1]:
            owner = getattr(owner, c)
        if p is not None:
            assert p.shape == table[key].shape, (key, p.shape, table[key].shape)
        object.__setattr__(owner, parts[-1], M.param(np.array(table[key], dtype=np.float64)))
        seen.add(key)
    assert seen == set(table), sorted(set(table) - seen)


def grads_of_link(link):
    return {path.lstrip("/"): p.grad for path, p in link.namedparams()}


def run(fn, float_inputs, ws):
    """fn(*Vars) -> Var or tuple of Vars; returns (outputs, input grads) after backward of sum_k sum(w_k * out_k)."""
    vs = [M.param(np.array(x, dtype=np.float64)) for x in float_inputs]
    outs = fn(*vs)
    outs = outs if isinstance(outs, tuple) else (outs,)
    total = None
    for o, w in zip(outs, ws):
        t = M.sum_(M.mul(o, M.const(w)))
        total = t if total is None else M.add(total, t)
    total.backward()
    return [o.data for o in outs], [v.grad for v in vs]


def check_and_save(name, params, ints, floats, ws, ref_out, ref_gin, ref_gp, ora_out, ora_gin, ora_gp, meta, skip_gin=False):
    def close(a, b, what):
        if a is None or b is None:
            assert a is None and b is None or (a is None and not np.any(b)) or (b is None and not np.any(a)), what
            return
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12, err_msg="%s: %s" % (name, what))
    for i, (a, b) in enumerate(zip(ref_out, ora_out)):
        close(a, b, "output %d" % i)
    if skip_gin:
        ref_gin = []
    for i, (a, b) in enumerate(zip(ref_gin, ora_gin)):
        close(a, b, "input gradient %d" % i)
    for k in params:
        close(ref_gp.get(k), ora_gp.get(k), "gradient of " + k)
    blob = {"meta": np.array(repr(meta))}
    blob.update({"param:" + k: v for k, v in params.items()})
    blob.update({"int:%d" % i: v for i, v in enumerate(ints)})
    blob.update({"float:%d" % i: v for i, v in enumerate(floats)})
    blob.update({"w:%d" % i: v for i, v in enumerate(ws)})
    blob.update({"out:%d" % i: v for i, v in enumerate(ref_out)})
    blob.update({"gin:%d" % i: v for i, v in enumerate(ref_gin) if v is not None})
    blob.update({"gparam:" + k: v for k, v in ref_gp.items() if v is not None})
    np.savez_compressed(os.path.join(HERE, "ref_%s.npz" % name), **blob)
    print("ref_%s.npz: %d outputs, %d parameter gradients, reference == oracle to 1e-10" % (name, len(ref_out), sum(v is not None for v in ref_gp.values())))


def ora_grads(table):
    return {k: v.grad for k, v in table.items()}


ACT = {"identity": CF.identity, "tanh": CF.tanh, "relu": CF.relu, "sigmoid": CF.sigmoid}


def main():
    from gcnbmp_synthetic import random_molecules
    rng = np.random.default_rng(20181018)
    # ---- modular GGNN (models/models/ggnn.py + models/update/ggnn_update.py + models/readout/ggnn_readout.py)
    for tag, tied, act, concat in (("ggnn_tied", True, "tanh", False), ("ggnn_untied", False, "identity", False), ("ggnn_concat", True, "identity", True)):
        H, O, T, mb, N = 12, 8, 3, 3, 9
        atoms, adj = random_molecules(rng, mb, N)
        adj = adj.astype(np.float64)
        params = R.init_params(R.ggnn_shapes(O, H, T, concat_hidden=concat, weight_tying=tied), rng, dtype=np.float64)
        w = rng.standard_normal((mb, O * (T if concat else 1)))
        net = ref_ggnn.GGNN(O, hidden_dim=H, n_layers=T, concat_hidden=concat, weight_tying=tied, activation=ACT[act])
        load_params(net, params)
        r_out, r_gin = run(lambda A: net(atoms, A), [adj], [w])
        tab = R.wrap_params(params)
        onet = R.GGNN(R.P(tab), O, H, T, concat_hidden=concat, weight_tying=tied, activation=act)
        o_out, o_gin = run(lambda A: onet(atoms, A), [adj], [w])
        check_and_save(tag, params, [atoms], [adj], [w], r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab),
                       dict(kind="ggnn", H=H, O=O, T=T, tied=tied, act=act, concat=concat))
    # ---- monolithic GGNN files: models/ggnn.py (what train_binary.py imports), models/ggnn_att.py (exposes the final atom
    # states for the co-attention), models/ggnn_dev.py (sum readout, per-step atom states)
    for tag, key, tied, sum_ro in (("mono_ggnn_py_tied", "ggnn_py", True, False), ("mono_ggnn_att_untied", "ggnn_att", False, False),
                                   ("mono_ggnn_att_tied", "ggnn_att", True, False), ("mono_ggnn_dev_sum", "ggnn_dev", True, True)):
        H, O, T, mb, N = 12, 8, 3, 3, 9
        atoms, adj = random_molecules(rng, mb, N)
        adj = adj.astype(np.float64)
        params = R.init_params(R.ggnn_mono_shapes(O, H, T, weight_tying=tied), rng, dtype=np.float64)
        ws = [rng.standard_normal((mb, H if sum_ro else O)), rng.standard_normal((mb, N, H))]
        net = ref_mono[key].GGNN(O, hidden_dim=H, n_layers=T, weight_tying=tied)
        load_params(net, params)
        with_atoms = hasattr(net, "get_atom_array")

        def both(n):
            def fn(A):
                g = n(atoms, A)
                return (g, n.get_atom_array()) if with_atoms else g
            return fn
        r_out, r_gin = run(both(net), [adj], ws)
        tab = R.wrap_params(params)
        onet = R.GGNNMono(R.P(tab), O, H, T, weight_tying=tied, sum_readout=sum_ro)
        o_out, o_gin = run(both(onet), [adj], ws)
        check_and_save(tag, params, [atoms], [adj], ws[:len(r_out)], r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab),
                       dict(kind="mono", H=H, O=O, T=T, tied=tied, sum_readout=sum_ro, with_atoms=with_atoms))
    # ---- the whole pair: train_binary.py:59-141 (encoder twice, co-attention, head) and the no-attention form
    # train_ddi_modify_eval2.py:50-104, with F.sigmoid_cross_entropy as the Classifier applies it (train_binary.py:524)
    PairAttn = load_class("train_binary.py", "GraphConvPredictorForPair")
    PairPlain = load_class("train_ddi_modify_eval2.py", "GraphConvPredictorForPair")
    PairEval3 = load_class("train_ddi_modify_eval3.py", "GraphConvPredictorForPair")     # co-attention on [h_first || h_last] atoms
    import cases
    for cname in ("C", "U", "A", "MU", "B", "CB", "E3", "L1", "E3B"):
        case = cases.pair_case(cname, seed=7)
        sp, params = case["spec"], case["params"]
        if sp["enc"] == "mono":
            mod = ref_mono["ggnn_dev"] if sp["sum_readout"] else ref_mono["ggnn_att"]
            enc = mod.GGNN(sp["O"], hidden_dim=sp["H"], n_layers=sp["T"], weight_tying=sp["tied"])
        elif sp["enc"] == "ggnn":
            enc = ref_ggnn.GGNN(sp["O"], hidden_dim=sp["H"], n_layers=sp["T"], weight_tying=sp["tied"], activation=ACT[sp.get("activation", "identity")])
        else:
            enc = ref_relgcn.RelGCN(out_channels=sp["O"], ch_list=list(sp["ch"]), scale_adj=sp["scale_adj"])
        d_atoms = sp["ch"][-1] if sp["enc"] == "relgcn" else sp["H"]
        if sp.get("first_last"):
            d_atoms = 2 * d_atoms
        attn = None
        if sp["attn"] == "nie":
            attn = ref_nie.NieFineCoattention(d_atoms, sp["O"], sp["head"], activation=CF.tanh)
        elif sp["attn"] == "vqa":
            attn = ref_vqa.VQAParallelCoattention(d_atoms, sp["O"], sp["head"])
        mlp = ref_hole.HolE(sp["K"], hidden_dims=sp["hole_hidden"])
        if sp.get("first_last"):
            net = PairEval3(enc, attn, mlp)
        else:
            net = PairAttn(enc, attn, mlp) if attn is not None else PairPlain(enc, mlp)
        load_params(net, params)
        a1, A1, a2, A2 = case["inputs"]
        logits = net(a1, A1, a2, A2)
        loss = CF.sigmoid_cross_entropy(logits, case["labels"])
        loss.backward()
        ref_gp = grads_of_link(net)
        o = cases.oracle_eval(case)
        np.testing.assert_allclose(logits.data, o["logits"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(loss.data, o["loss"], rtol=1e-10)
        for k in params:
            if ref_gp[k] is None or o["grads"][k] is None:      # e.g. the readout when the co-attention ignores g_1, g_2
                assert (ref_gp[k] is None or not np.any(ref_gp[k])) and (o["grads"][k] is None or not np.any(o["grads"][k])), k
                continue
            np.testing.assert_allclose(ref_gp[k], o["grads"][k], rtol=1e-9, atol=1e-13, err_msg="pair %s: %s" % (cname, k))
        blob = {"meta": np.array(repr(dict(kind="pair", case=cname, seed=7))), "logits": logits.data, "loss": np.asarray(loss.data)}
        blob.update({"gparam:" + k: v for k, v in ref_gp.items() if v is not None})
        np.savez_compressed(os.path.join(HERE, "ref_pair_%s.npz" % cname), **blob)
        print("ref_pair_%s.npz: logits %s, loss %.6f, %d parameter gradients, reference == oracle to 1e-10" % (cname, logits.shape, float(loss.data), len(ref_gp)))
    # ---- GIN (models/gin.py; dropout off): tied (ONE update step, the loop bound at :154) and untied
    for tag, tied, concat in (("gin_tied", True, False), ("gin_untied_concat", False, True)):
        H, O, T, mb, N = 12, 8, 3, 3, 9
        atoms, adj = random_molecules(rng, mb, N)
        adj = adj.astype(np.float64)
        params = R.init_params(R.gin_shapes(O, H, T, concat_hidden=concat, weight_tying=tied), rng, dtype=np.float64)
        w = rng.standard_normal((mb, O * (T if concat else 1)))
        net = ref_gin.GIN(O, hidden_dim=H, n_layers=T, dropout_ratio=0.0, concat_hidden=concat, weight_tying=tied, activation=CF.tanh)
        load_params(net, params)
        r_out, r_gin = run(lambda: net(atoms, adj), [], [w])
        tab = R.wrap_params(params)
        onet = R.GIN(R.P(tab), O, H, T, concat_hidden=concat, weight_tying=tied, activation="tanh")
        o_out, o_gin = run(lambda: onet(atoms, adj), [], [w])
        check_and_save(tag, params, [atoms], [adj], [w], r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab),
                       dict(kind="gin", H=H, O=O, T=T, tied=tied, concat=concat, act="tanh"))
    # ---- NFP (models/models/nfp.py): (mb, N, N) adjacency with self connections, as the NFP preprocessor builds it
    nfp_rng = np.random.default_rng(20191105)          # its own stream: the fixtures above and below keep their bits
    for tag, mb, N in (("nfp", 4, 9), ("nfp_ragged", 3, 13)):
        H, O, T = 12, 8, 3
        atoms, adj4 = random_molecules(nfp_rng, mb, N)
        adj = adj4.sum(axis=1).astype(np.float64)
        adj = adj + np.eye(N)[None] * (atoms > 0)[:, :, None]      # self connection on the real atoms; padded atoms keep degree 0
        if tag == "nfp_ragged":
            adj[0, 1, 2] = adj[0, 2, 1] = 0.0                       # degrees need not be symmetric-consistent: column sums rule
        params = R.init_params(R.nfp_shapes(O, H, T), nfp_rng, dtype=np.float64)
        ws = [nfp_rng.standard_normal((mb, O)), nfp_rng.standard_normal((mb, N, H))]
        net = ref_nfp.NFP(O, hidden_dim=H, n_layers=T)
        load_params(net, params)

        def both(n):
            def fn():
                g = n(atoms, adj)
                return g, n.get_atom_array()
            return fn
        r_out, r_gin = run(both(net), [], ws)
        tab = R.wrap_params(params)
        o_out, o_gin = run(both(R.NFP(R.P(tab), O, H, T)), [], ws)
        check_and_save(tag, params, [atoms], [adj], ws, r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab),
                       dict(kind="nfp", H=H, O=O, T=T))
    # ---- BiMPM (models/coattention/bimpm.py): all three matchings; N1 != N2; outputs have 3 * head columns
    bi_rng = np.random.default_rng(20190516)
    for tag, mb, N1, N2, H, head in (("bimpm", 3, 6, 9, 12, 4), ("bimpm_wide", 2, 11, 7, 20, 5)):
        a1, a2 = bi_rng.standard_normal((mb, N1, H)) * 0.7, bi_rng.standard_normal((mb, N2, H)) * 0.7
        params = R.init_params(R.bimpm_shapes(H, head), bi_rng, dtype=np.float64)
        ws = [bi_rng.standard_normal((mb, 3 * head)), bi_rng.standard_normal((mb, 3 * head))]
        net = ref_bimpm.BiMPM(H, 8, head)
        load_params(net, params)
        r_out, r_gin = run(lambda x1, x2: net(x1, None, x2, None), [a1, a2], ws)
        tab = R.wrap_params(params)
        onet = R.BiMPM(R.P(tab), H, 8, head)
        o_out, o_gin = run(lambda x1, x2: onet(x1, None, x2, None), [a1, a2], ws)
        check_and_save(tag, params, [], [a1, a2], ws, r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab),
                       dict(kind="coattn_bimpm", H=H, O=8, head=head))
    # ---- bare GGNNUpdate with state threading (two calls, then reset, then one call)
    H, mb, N = 8, 2, 7
    _, adj = random_molecules(rng, mb, N)
    adj = adj.astype(np.float64)
    h_in = rng.standard_normal((mb, N, H)) * 0.5
    params = R.init_params({k[len("update_layers/0/"):]: v for k, v in R.ggnn_shapes(H, H, 1).items() if k.startswith("update_layers/0/")}, rng, dtype=np.float64)
    w = rng.standard_normal((mb, N, H))
    link = ref_ggnn_update.GGNNUpdate(hidden_dim=H)
    load_params(link, params)

    def thread(l):
        def fn(h, A):
            l.reset_state()
            h1 = l(h, A)
            h2 = l(h1, A)
            return h2
        return fn
    r_out, r_gin = run(thread(link), [h_in, adj], [w])
    tab = R.wrap_params(params)
    o_out, o_gin = run(thread(R.GGNNUpdate(R.P(tab), H)), [h_in, adj], [w])
    check_and_save("ggnn_update_threaded", params, [], [h_in, adj], [w], r_out, r_gin, grads_of_link(link), o_out, o_gin, ora_grads(tab), dict(kind="ggnn_update", H=H))
    # ---- RelGCN (models/relgcn.py + models/update/relgcn_update.py)
    for tag, scale in (("relgcn_scaled", True), ("relgcn_plain", False)):
        ch, O, mb, N = [8, 12, 16], 8, 3, 9
        atoms, adj = random_molecules(rng, mb, N)
        adj = adj.astype(np.float64)
        params = R.init_params(R.relgcn_shapes(O, ch), rng, dtype=np.float64)
        w = rng.standard_normal((mb, O))
        net = ref_relgcn.RelGCN(out_channels=O, ch_list=list(ch), scale_adj=scale)
        load_params(net, params)
        r_out, r_gin = run(lambda: net(atoms, adj), [], [w])          # the adjacency is data (rescale_adj: relgcn.py:18-28)
        tab = R.wrap_params(params)
        onet = R.RelGCN(R.P(tab), O, ch_list=list(ch), scale_adj=scale)
        o_out, o_gin = run(lambda: onet(atoms, adj), [], [w])
        check_and_save(tag, params, [atoms, ], [adj], [w], r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab), dict(kind="relgcn", ch=ch, O=O, scale=scale))
    # ---- fine co-attention (nie / vqa / pooling)
    for tag, mk_ref, mk_ora, head in (
            ("coattn_nie", lambda H, O, hd: ref_nie.NieFineCoattention(H, O, hd, activation=CF.tanh), lambda p, H, O, hd: R.NieFineCoattention(p, H, O, hd, activation="tanh"), 4),
            ("coattn_vqa", lambda H, O, hd: ref_vqa.VQAParallelCoattention(H, O, hd), lambda p, H, O, hd: R.VQAParallelCoattention(p, H, O, hd), 3),
            ("coattn_fourier", lambda H, O, hd: ref_nie.FourierFineCoattention(H, O, hd, activation=CF.tanh), lambda p, H, O, hd: R.FourierFineCoattention(p, H, O, hd, activation="tanh"), 4),
            ("coattn_pool", lambda H, O, hd: ref_pool.PoolingFineCoattention(H, O), lambda p, H, O, hd: R.PoolingFineCoattention(p, H, O), None)):
        H, O, mb, N1, N2 = 12, 8, 3, 6, 9
        a1, a2 = rng.standard_normal((mb, N1, H)) * 0.5, rng.standard_normal((mb, N2, H)) * 0.5
        g1, g2 = rng.standard_normal((mb, O)), rng.standard_normal((mb, O))
        params = R.init_params(R.coattn_shapes(H, O, head), rng, dtype=np.float64)
        ws = [rng.standard_normal((mb, O)), rng.standard_normal((mb, O))]
        net = mk_ref(H, O, head)
        load_params(net, params)
        r_out, r_gin = run(lambda x1, x2: net(x1, M.const(g1), x2, M.const(g2)), [a1, a2], ws)
        tab = R.wrap_params(params)
        onet = mk_ora(R.P(tab), H, O, head)
        o_out, o_gin = run(lambda x1, x2: onet(x1, M.const(g1), x2, M.const(g2)), [a1, a2], ws)
        check_and_save(tag, params, [], [a1, a2], ws, r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab), dict(kind=tag, H=H, O=O, head=head))
    # ---- vector-query co-attentions: alternating / parallel / circular (the graph vectors g_1, g_2 are live inputs here)
    for tag, kind, mk_ref, mk_ora, head in (
            ("coattn_alter", "alter", lambda H, O, hd: ref_alter.AlternatingCoattention(H, O, hd, weight_tying=True), lambda p, H, O, hd: R.AlternatingCoattention(p, H, O, hd), 4),
            ("coattn_para", "para", lambda H, O, hd: ref_para.ParallelCoattention(H, O, hd, activation=CF.tanh, weight_tying=True), lambda p, H, O, hd: R.ParallelCoattention(p, H, O, hd), 1),
            ("coattn_circ", "circ", lambda H, O, hd: ref_para.CircularParallelCoattention(H, O, activation=CF.tanh), lambda p, H, O, hd: R.CircularParallelCoattention(p, H, O), 1),
            ("coattn_global", "global", lambda H, O, hd: ref_global.GlobalCoattention(H, O, weight_tying=True), lambda p, H, O, hd: R.GlobalCoattention(p, H, O), 1),
            ("coattn_neural", "neural", lambda H, O, hd: ref_neural.NeuralCoattention(H, O, activation=CF.tanh, weight_tying=True), lambda p, H, O, hd: R.NeuralCoattention(p, H, O, activation="tanh"), 1)):
        H, O, mb, N1, N2 = 12, 8, 3, 6, 9
        a1, a2 = rng.standard_normal((mb, N1, H)) * 0.5, rng.standard_normal((mb, N2, H)) * 0.5
        g1, g2 = rng.standard_normal((mb, O)) * 0.5, rng.standard_normal((mb, O)) * 0.5
        params = R.init_params(R.vector_coattn_shapes(kind, H, O, head), rng, dtype=np.float64)
        ws = [rng.standard_normal((mb, O)), rng.standard_normal((mb, O))]
        net = mk_ref(H, O, head)
        load_params(net, params)
        r_out, r_gin = run(lambda x1, x2, q1, q2: net(x1, q1, x2, q2), [a1, a2, g1, g2], ws)
        tab = R.wrap_params(params)
        onet = mk_ora(R.P(tab), H, O, head)
        o_out, o_gin = run(lambda x1, x2, q1, q2: onet(x1, q1, x2, q2), [a1, a2, g1, g2], ws)
        check_and_save(tag, params, [], [a1, a2, g1, g2], ws, r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab), dict(kind=tag, H=H, O=O, head=head))
    # ---- the Deep / VeryDeep / ExtremeDeep Nie variants (energy on the original atoms, heads on the transformed ones)
    for tag, rcls, ocls, nl in (("coattn_deep", ref_nie.DeepNieFineCoattention, R.DeepNieFineCoattention, 1),
                                ("coattn_very_deep", ref_nie.VeryDeepNieFineCoattention, R.VeryDeepNieFineCoattention, 2),
                                ("coattn_extreme_deep", ref_nie.ExtremeDeepNieFineCoattention, R.ExtremeDeepNieFineCoattention, 3)):
        H, O, mb, N1, N2, head = 12, 8, 3, 6, 9, 4
        a1, a2 = rng.standard_normal((mb, N1, H)) * 0.5, rng.standard_normal((mb, N2, H)) * 0.5
        params = R.init_params(R.deep_coattn_shapes(H, O, head, nl), rng, dtype=np.float64)
        ws = [rng.standard_normal((mb, O)), rng.standard_normal((mb, O))]
        net = rcls(H, O, head, activation=CF.tanh)
        load_params(net, params)
        r_out, r_gin = run(lambda x1, x2: net(x1, None, x2, None), [a1, a2], ws)
        tab = R.wrap_params(params)
        onet = ocls(R.P(tab), H, O, head, activation="tanh")
        o_out, o_gin = run(lambda x1, x2: onet(x1, None, x2, None), [a1, a2], ws)
        check_and_save(tag, params, [], [a1, a2], ws, r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab), dict(kind=tag, H=H, O=O, head=head))
    # ---- heads: HolE (both files), MLP, SymMLP, NTN, DistMult
    D, K, mb = 12, 3, 5
    l, r = rng.standard_normal((mb, D)), rng.standard_normal((mb, D))
    w = rng.standard_normal((mb, K))
    heads = [
        ("head_hole", lambda: ref_hole.HolE(K, hidden_dims=(8,)), R.hole_shapes(D, K, (8,)), lambda p: R.HolE(p, K, hidden_dims=(8,)), False),
        ("head_hole_mlp_py", lambda: ref_mlp.HolE(K, hidden_dims=()), R.hole_shapes(D, K, (), layers_name="layers"), lambda p: R.HolE(p, K, hidden_dims=(), layers_name="layers"), False),
        ("head_symmlp", lambda: ref_mlp.SymMLP(K, hidden_dims=(8,)), R.head_shapes("symmlp", D, K, (8,)), lambda p: R.SymMLP(p, K, (8,)), False),
        ("head_ntn", lambda: ref_mlp.NTN(D, D, K, ntn_out_dim=4, hidden_dims=(6,)), R.head_shapes("ntn", D, K, (6,), mid=4), lambda p: R.NTN(p, D, D, K, 4, (6,)), False),
        ("head_distmult", lambda: ref_mlp.DistMult(D, D, K, dm_out_dim=4, hidden_dims=(6,)), R.head_shapes("distmult", D, K, (6,), mid=4), lambda p: R.DistMult(p, D, D, K, 4, (6,)), False),
        ("head_mlp", lambda: ref_mlp.MLP(K, hidden_dims=(8, 6)), R.head_shapes("mlp", D, K, (8, 6)), lambda p: R.MLP(p, K, (8, 6)), True),
    ]
    for tag, mk_ref, shapes, mk_ora, single in heads:
        params = R.init_params(shapes, rng, dtype=np.float64)
        net = mk_ref()
        load_params(net, params)
        call = (lambda n: (lambda a, b: n(CF.concat((a, b), axis=-1)))) if single else (lambda n: (lambda a, b: n(a, b)))
        r_out, r_gin = run(call(net), [l, r], [w])
        tab = R.wrap_params(params)
        o_out, o_gin = run(call(mk_ora(R.P(tab))), [l, r], [w])
        # SymMLP: models/mlp.py:105 builds its input with xp.concatenate on the Variables, which leaves the graph -- the reference
        # defines no input gradient there (ours keeps it flowing); outputs and parameter gradients are compared
        check_and_save(tag, params, [], [l, r], [w], r_out, r_gin, grads_of_link(net), o_out, o_gin, ora_grads(tab), dict(kind=tag, D=D, K=K),
                       skip_gin=(tag == "head_symmlp"))
    # ---- GGNNReadout variants (models/readout/ggnn_readout.py)
    H, O, mb, N = 8, 12, 3, 7
    h, h0 = rng.standard_normal((mb, N, H)) * 0.5, rng.standard_normal((mb, N, H)) * 0.5
    mask = (rng.random((mb, N)) < 0.7).astype(np.float64)
    for tag, use_h0, use_mask, nobias, act, agg in (("readout_full", True, True, False, "tanh", "tanh"), ("readout_noh0_nobias", False, False, True, "identity", "sigmoid")):
        kin = 2 * H if use_h0 else H
        shapes = {"i_layer/W": (O, kin), "j_layer/W": (O, kin)}
        if not nobias:
            shapes.update({"i_layer/b": (O,), "j_layer/b": (O,)})
        params = R.init_params(shapes, rng, dtype=np.float64)
        w = rng.standard_normal((mb, O))
        net = ref_ggnn_readout.GGNNReadout(O, hidden_dim=H, nobias=nobias, activation=ACT[act], activation_agg=ACT[agg])
        load_params(net, params)
        f = lambda n: (lambda x, x0: n(x, x0 if use_h0 else None, mask if use_mask else None))
        r_out, r_gin = run(f(net), [h, h0], [w])
        tab = R.wrap_params(params)
        o_out, o_gin = run(f(R.GGNNReadout(R.P(tab), O, H, nobias=nobias, activation=act, activation_agg=agg)), [h, h0], [w])
        check_and_save(tag, params, [], [h, h0] + ([mask] if use_mask else []), [w], r_out, r_gin[:2], grads_of_link(net), o_out, o_gin[:2], ora_grads(tab),
                       dict(kind="readout", H=H, O=O, use_h0=use_h0, use_mask=use_mask, nobias=nobias, act=act, agg=agg))


if __name__ == "__main__":
    sys.modules["gcnbmp_synthetic"] = importlib.util.module_from_spec(
        importlib.util.spec_from_file_location("gcnbmp_synthetic", os.path.join(ROOT, "gcn-bmp_b200", "gcnbmp", "synthetic.py")))
    sys.modules["gcnbmp_synthetic"].__spec__.loader.exec_module(sys.modules["gcnbmp_synthetic"])
    main()
