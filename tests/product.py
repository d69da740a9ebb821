"""Build the product-side (gcnbmp) model for a parity case and run it on the GPU."""
import numpy as np
import torch

import gcnbmp


def product_model(spec, params):
    f = gcnbmp.functions
    if spec["enc"] == "mono":
        enc = gcnbmp.GGNNMono(spec["O"], spec["H"], spec["T"], weight_tying=spec["tied"],
                              sum_readout=spec["sum_readout"])
        d_atoms = spec["H"]
    elif spec["enc"] == "ggnn":
        enc = gcnbmp.GGNN(spec["O"], spec["H"], spec["T"], weight_tying=spec["tied"],
                          activation=getattr(f, spec.get("activation", "identity")))
        d_atoms = spec["H"]
    else:
        enc = gcnbmp.RelGCN(spec["O"], ch_list=spec["ch"], scale_adj=spec["scale_adj"])
        d_atoms = spec["ch"][-1]
    attn = None
    if spec.get("first_last"):
        d_atoms = 2 * d_atoms
    if spec["attn"] == "nie":
        attn = gcnbmp.NieFineCoattention(d_atoms, spec["O"], spec["head"], activation=f.tanh)
    elif spec["attn"] == "vqa":
        attn = gcnbmp.VQAParallelCoattention(d_atoms, spec["O"], spec["head"])
    elif spec["attn"] == "pool":
        attn = gcnbmp.PoolingFineCoattention(d_atoms, spec["O"])
    mlp = gcnbmp.HolE(spec["K"], hidden_dims=spec["hole_hidden"])
    model = gcnbmp.GraphConvPredictorForPair(enc, attn, mlp, first_last_atoms=bool(spec.get("first_last")))
    model.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
    return model


def product_eval(case, mode=None):
    model = product_model(case["spec"], case["params"])
    if mode is not None:
        model.graph_conv.mode = mode
        if "attn" in model._children:
            model.attn.mode = mode
    a1, A1, a2, A2 = case["inputs"]
    model.cleargrads()
    logits = model(a1, A1.astype(np.float32), a2, A2.astype(np.float32))
    loss = gcnbmp.sigmoid_cross_entropy(logits, case["labels"])
    loss.backward()
    torch.cuda.synchronize()
    return dict(loss=loss.item(), logits=logits.detach().cpu().numpy(), grads=model.grad_dict(), model=model)


def rel_err(a, b, floor=1e-30):
    """max |a-b| / max |b| -- the '1e-4 relative' of BASELINE.json north_star.  `floor` bounds the
    denominator for quantities that are exactly zero by symmetry (e.g. the gradient of the
    bilinear bias under an identity activation: both softmaxes are shift-invariant)."""
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), floor))
