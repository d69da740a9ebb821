"""CPU tests of the host-side mirror: Chainer parameter names/shapes, flat buffers, the synthetic
generator, the FLOP model, golden-vector pins of the oracle, and the 2-rank data-parallel
reduction on gloo."""
import os
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import reference_path as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["A", "C", "U", "M", "MU", "B"])
def test_oracle_reproduces_golden_vectors(name):
    case = cases.pair_case(name, seed=2018)
    o = cases.oracle_eval(case)
    g = np.load(os.path.join(GOLD, "pair_%s.npz" % name))
    np.testing.assert_allclose(o["logits"], g["logits"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(float(o["loss"]), float(g["loss"]), rtol=1e-9)
    for k, v in o["grads"].items():
        ref = g["grad:" + k]
        np.testing.assert_allclose(v, ref, rtol=2e-6, atol=2e-7 * max(np.abs(ref).max(), 1e-30) + 1e-12, err_msg=k)


def _names(link):
    return {k.lstrip("/"): tuple(p.shape) for k, p in link.namedparams()}


def test_link_parameter_names_match_chainer_paths():
    import gcnbmp
    assert _names(gcnbmp.GGNN(24, 32, 3)) == R.ggnn_shapes(24, 32, 3)
    assert _names(gcnbmp.GGNN(24, 32, 3, weight_tying=False, concat_hidden=True)) == \
        R.ggnn_shapes(24, 32, 3, weight_tying=False, concat_hidden=True)
    assert _names(gcnbmp.GGNNMono(24, 32, 3, weight_tying=False)) == R.ggnn_mono_shapes(24, 32, 3, weight_tying=False)
    assert _names(gcnbmp.RelGCN(16, ch_list=[16, 32, 8])) == R.relgcn_shapes(16, [16, 32, 8])
    assert _names(gcnbmp.NieFineCoattention(32, 16, 8)) == R.coattn_shapes(32, 16, 8)
    assert _names(gcnbmp.PoolingFineCoattention(32, 16)) == R.coattn_shapes(32, 16, None)
    hole = gcnbmp.HolE(5, hidden_dims=(12, 6))
    hole.load_params(R.init_params(R.hole_shapes(16, 5, (12, 6)), np.random.default_rng(0)))
    assert _names(hole) == R.hole_shapes(16, 5, (12, 6))
    assert gcnbmp.GGNNUpdate().hidden_dim == 16 and gcnbmp.GGNN(8).n_layers == 4     # reference defaults


def test_load_params_rejects_wrong_shapes_and_missing_lazy_params():
    import gcnbmp
    with pytest.raises(ValueError):
        gcnbmp.GGNNUpdate(16).load_params({"graph_linear/W": np.zeros((3, 3), np.float32)})
    with pytest.raises(KeyError):
        gcnbmp.HolE(1, hidden_dims=(4,)).load_params({})


def test_flat_buffers_are_aligned_views():
    import gcnbmp
    m = gcnbmp.GraphConvPredictorForPair(gcnbmp.GGNNMono(16, 16, 2), gcnbmp.NieFineCoattention(16, 16, 8),
                                          gcnbmp.HolE(1, ()))
    m.mlp.l_out.ensure(16)
    before = {k: p.detach().clone() for k, p in m.namedparams()}
    flat, gflat = m.flatten_parameters()
    for k, p in m.namedparams():
        assert torch.equal(p.detach(), before[k])
        assert (p.data_ptr() - flat.data_ptr()) % 256 == 0 and (p.grad.data_ptr() - gflat.data_ptr()) % 256 == 0
        assert p.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    with torch.no_grad():
        flat.zero_()
    assert all(float(p.abs().sum()) == 0 for p in m.params())


def test_synthetic_molecules_layout():
    from gcnbmp import synthetic
    a, A = synthetic.random_molecules(np.random.default_rng(1), 64, 50)
    b, B = synthetic.random_molecules(np.random.default_rng(1), 64, 50)
    assert np.array_equal(a, b) and np.array_equal(A, B)                 # deterministic
    assert a.dtype == np.int32 and A.dtype == np.float32 and A.shape == (64, 4, 50, 50)
    assert np.array_equal(A, A.transpose(0, 1, 3, 2))                    # symmetric
    assert set(np.unique(A)) <= {0.0, 1.0}
    assert A[:, :, np.arange(50), np.arange(50)].sum() == 0              # no self loops
    assert A.sum(axis=(1, 3)).max() <= 4                                 # degree cap
    n = (a != 0).sum(axis=1)
    assert n.min() >= 25 and n.max() <= 50
    pad = a == 0
    assert A.sum(axis=1)[pad].sum() == 0                                 # padded atoms have no bonds


def test_flop_model_matches_baseline_md():
    from gcnbmp.train import algorithmic_flops
    assert abs(algorithmic_flops(128, 6, 64, 4, 128, 8, 86)["pair_fwd"] / 1e6 - 385.3) < 0.1   # config C
    d = algorithmic_flops(256, 8, 64, 4, 256, 8, 1, readout="r2", attn=False)
    assert abs(d["pair_fwd"] / 1e6 - 1879.2) < 0.5                                             # config D


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from gcnbmp import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = cases.pair_case("U", seed=5)
    n = case["labels"].shape[0]
    s, e = parallel.shard_bounds(n, rank, world)
    table = R.wrap_params(case["params"])
    model = cases.oracle_model(case["spec"], table)
    from oracle import minichainer as F
    logits = model(*[x[s:e] for x in case["inputs"]])
    y = case["labels"][s:e]
    # shard loss normalised by the GLOBAL count (what PairTrainer passes as global_count)
    count = float((case["labels"] != -1).sum())
    local_cnt = max(int((y != -1).sum()), 1)
    loss = F.sigmoid_cross_entropy(logits, y)
    loss.backward(seed=np.asarray(local_cnt / count))
    flat = torch.cat([torch.from_numpy(np.asarray(table[k].grad if table[k].grad is not None
                                                 else np.zeros_like(table[k].data)).reshape(-1)) for k in sorted(table)])
    parallel.allreduce_sum_(flat)
    if rank == 0:
        np.save(out, flat.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_rank(tmp_path):
    """world_size 2 on gloo: shard the pairs, normalise by the global count, SUM-allreduce the flat
    gradient -> identical to the full-batch gradient (the N-GPU == 1-GPU contract of SURVEY 8e)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "flat.npy")
    port = 29500 + os.getpid() % 2000
    mp.start_processes(_dp_worker, args=(2, port, out), nprocs=2, join=True, start_method="spawn")
    got = np.load(out)
    case = cases.pair_case("U", seed=5)
    full = cases.oracle_eval(case)["grads"]
    ref = np.concatenate([full[k].reshape(-1) for k in sorted(full)])
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)


def test_shard_bounds_cover_everything():
    from gcnbmp import parallel
    for n in (1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_flatten_parameters_refuses_uninitialised_lazy_layers():
    """A lazily-shaped layer created after flattening would sit outside the flat buffers (no allreduce, no Adam update)."""
    import gcnbmp
    enc = gcnbmp.GGNNMono(16, 16, 2)
    head = gcnbmp.HolE(1, hidden_dims=())            # l_out is shaped at its first call
    model = gcnbmp.GraphConvPredictorForPair(enc, None, head)
    with pytest.raises(ValueError, match="lazily-shaped"):
        model.flatten_parameters()
    head.l_out.ensure(16)
    flat, gflat = model.flatten_parameters()
    assert flat.numel() == gflat.numel() > 0


def test_load_params_after_flattening_keeps_the_flat_views():
    """Loading a snapshot into a flattened model must write THROUGH the views (the optimiser updates the flat buffer)."""
    import gcnbmp
    enc = gcnbmp.GGNNMono(16, 16, 2)
    head = gcnbmp.HolE(1, hidden_dims=())
    head.l_out.ensure(16)
    model = gcnbmp.GraphConvPredictorForPair(enc, None, head)
    flat, _ = model.flatten_parameters()
    snap = {k: v + 1.0 for k, v in model.param_dict().items()}
    before = flat.clone()
    model.load_params(snap)
    n_real = sum(v.size for v in snap.values())
    assert abs(float((flat - before).sum()) - n_real) < 1e-2 * n_real          # every real element moved by +1 inside the flat buffer
    for k, p in model.namedparams():
        assert p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + flat.numel() * 4, k


def test_cleargrads_keeps_flat_gradient_views():
    import gcnbmp
    enc = gcnbmp.GGNNMono(16, 16, 2)
    head = gcnbmp.HolE(1, hidden_dims=())
    head.l_out.ensure(16)
    model = gcnbmp.GraphConvPredictorForPair(enc, None, head)
    _, gflat = model.flatten_parameters()
    gflat.fill_(3.0)
    model.cleargrads()
    for k, p in model.namedparams():            # (the alignment padding between the views is not anybody's gradient)
        assert p.grad is not None and p.grad._base is not None, k
        assert float(p.grad.abs().sum()) == 0.0, k
    assert float(gflat.abs().sum()) > 0.0
    plain = gcnbmp.GGNNMono(16, 16, 2)              # unflattened model: chainer semantics (gradients dropped)
    for p in plain.params():
        p.grad = torch.ones_like(p)
    plain.cleargrads()
    assert all(p.grad is None for p in plain.params())
