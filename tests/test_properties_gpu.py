"""Size-independent properties at BASELINE.json's FULL shapes (where the fp64 oracle is too slow to be the checker):
  * atom relabelling: permuting the atoms of every molecule (ids, adjacency rows and columns alike) leaves graph vectors,
    pooled co-attention vectors and logits unchanged, and permutes the atom states -- message passing, readout and the
    fine-grained co-attention are permutation-equivariant in the reference's math;
  * batch independence: a pair's logits do not depend on which other pairs share the launch (persistent-CTA tiling, tile pairing
    of molecules, micro-batching), bit for bit;
  * gradient linearity: the parameter gradient of a batch is the sum of the gradients of its halves (what the trainer's
    micro-batching and the data-parallel allreduce rely on).
Config C/E: GGNN H128 T6 + Nie co-attention (head 8) + HolE -> 86;  config D: GGNN H256 T8 + R1 readout + HolE -> 1;
config B: RelGCN 64 x 4.  N = 64 padded."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pairs(rng, mb, N=64):
    from gcnbmp import synthetic
    a1, A1 = synthetic.random_molecules(rng, mb, N, pad_to=N)
    a2, A2 = synthetic.random_molecules(rng, mb, N, pad_to=N)
    return [torch.tensor(x).cuda() for x in (a1, A1, a2, A2)]


def _permute(atoms, adj, perm):
    return atoms[:, perm].contiguous(), adj[:, :, perm][:, :, :, perm].contiguous()


def _model(cfg):
    import gcnbmp
    f = gcnbmp.functions
    if cfg == "C":
        enc, attn, head = gcnbmp.GGNNMono(128, 128, 6), gcnbmp.NieFineCoattention(128, 128, 8, activation=f.tanh), gcnbmp.HolE(86, hidden_dims=())
    elif cfg == "D":
        enc, attn, head = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8), None, gcnbmp.HolE(1, hidden_dims=())
    else:
        enc, attn, head = gcnbmp.RelGCN(64, ch_list=[64, 64, 64, 64, 64], scale_adj=True), None, gcnbmp.HolE(1, hidden_dims=())
    return gcnbmp.GraphConvPredictorForPair(enc, attn, head)


def _set_mode(model, mode):
    model.graph_conv.mode = mode
    if model.attn is not None:
        model.attn.mode = mode


@pytest.mark.parametrize("cfg,mode,mb,tol", [("C", "f32", 64, 2e-4), ("C", "bf16", 1024, 3e-2), ("D", "f32", 16, 2e-4), ("D", "bf16", 1024, 3e-2),
                                             ("B", "f32", 256, 2e-4), ("B", "bf16", 1024, 3e-2)])
def test_atom_relabelling_invariance_at_full_shapes(cfg, mode, mb, tol):
    import gcnbmp
    rng = np.random.default_rng(7)
    a1, A1, a2, A2 = _pairs(rng, mb)
    model = _model(cfg)
    _set_mode(model, gcnbmp.MODE_BF16 if mode == "bf16" else gcnbmp.MODE_F32)
    p1, p2 = torch.tensor(rng.permutation(64)).cuda(), torch.tensor(rng.permutation(64)).cuda()
    with torch.no_grad():
        ref = model(a1, A1, a2, A2)
        atoms_ref = model.graph_conv.get_atom_array().clone()          # states of the second molecule
        b1, B1 = _permute(a1, A1, p1)
        b2, B2 = _permute(a2, A2, p2)
        got = model(b1, B1, b2, B2)
        atoms_got = model.graph_conv.get_atom_array()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol * scale, float((got - ref).abs().max()) / scale
    # equivariance of the atom states: row p2[i] of the reference is row i of the permuted run
    err = float((atoms_got - atoms_ref[:, p2]).abs().max()) / float(atoms_ref.abs().max())
    assert err <= tol, err


@pytest.mark.parametrize("cfg,mode", [("C", "bf16"), ("D", "bf16"), ("B", "bf16"), ("C", "f32")])
def test_logits_do_not_depend_on_batch_composition(cfg, mode):
    import gcnbmp
    rng = np.random.default_rng(11)
    mb = 1500 if mode == "bf16" else 96
    a1, A1, a2, A2 = _pairs(rng, mb)
    model = _model(cfg)
    _set_mode(model, gcnbmp.MODE_BF16 if mode == "bf16" else gcnbmp.MODE_F32)
    with torch.no_grad():
        whole = model(a1, A1, a2, A2)
        cut = 2 * (mb // 3) + 1                # odd cut: the molecules pair up differently inside the two-molecule tiles
        parts = torch.cat([model(a1[:cut], A1[:cut], a2[:cut], A2[:cut]), model(a1[cut:], A1[cut:], a2[cut:], A2[cut:])])
        rev = model(a1.flip(0).contiguous(), A1.flip(0).contiguous(), a2.flip(0).contiguous(), A2.flip(0).contiguous()).flip(0)
    assert torch.equal(whole, parts)
    assert torch.equal(whole, rev)


@pytest.mark.parametrize("cfg,mode,mb", [("C", "bf16", 512), ("C", "f32", 32), ("B", "bf16", 512)])
def test_batch_gradient_is_the_sum_of_its_halves(cfg, mode, mb):
    import gcnbmp
    rng = np.random.default_rng(13)
    a1, A1, a2, A2 = _pairs(rng, mb)
    K = 86 if cfg == "C" else 1
    y = torch.tensor((rng.random((mb, K)) < 0.2).astype(np.int32)).cuda()
    model = _model(cfg)
    _set_mode(model, gcnbmp.MODE_BF16 if mode == "bf16" else gcnbmp.MODE_F32)
    with torch.no_grad():
        model(a1[:2], A1[:2], a2[:2], A2[:2])                          # materialise lazily-shaped layers
    count = float(mb * K)

    def grads(sl):
        model.cleargrads()
        gcnbmp.sigmoid_cross_entropy(model(a1[sl], A1[sl], a2[sl], A2[sl]), y[sl], count=count).backward()
        return {k: np.asarray(v, np.float64).copy() for k, v in model.grad_dict().items() if v is not None}

    whole = grads(slice(0, mb))
    h1, h2 = grads(slice(0, mb // 2)), grads(slice(mb // 2, mb))
    for k, g in whole.items():
        s = h1[k] + h2[k]
        denom = max(np.abs(g).max(), 1e-12)
        assert np.abs(s - g).max() / denom <= 2e-3, (k, np.abs(s - g).max() / denom)      # fp32 atomics: order-dependent rounding only
