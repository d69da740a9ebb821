"""MODE_F32 encoder: tensor-core (bf16 hi/lo split, csrc/ggnn_x3.cu) vs FFMA (csrc/ggnn.cu) -- outputs, gradients, time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, gcnbmp
from gcnbmp import functional as F, synthetic


def run(H, T, N, mb, tied, x3, n_max=None, time_it=False):
    torch.manual_seed(0)
    np.random.seed(0)
    F.F32_TENSOR_CORES = x3
    rng = np.random.default_rng(5)
    atoms, adj = synthetic.random_molecules(rng, mb, n_max or N, pad_to=N)
    enc = gcnbmp.GGNN(out_dim=min(H, 64), hidden_dim=H, n_layers=T, weight_tying=tied)
    rs = np.random.default_rng(7)
    for k, p in sorted(enc.namedparams()):
        p.data.copy_(torch.tensor(rs.standard_normal(tuple(p.shape)) * (0.5 / np.sqrt(max(p.shape[-1], 1))), dtype=torch.float32))
    a_t = torch.tensor(atoms, device="cuda")
    A_t = torch.tensor(adj, device="cuda")
    g = enc(a_t, A_t)
    at = enc.get_atom_array()
    w = torch.tensor(rs.standard_normal(tuple(at.shape)), dtype=torch.float32, device="cuda")
    loss = (at * w).sum() + g.sum()
    loss.backward()
    out = {"g": g.detach().cpu().numpy(), "atoms": at.detach().cpu().numpy()}
    for k, p in sorted(enc.namedparams()):
        out["d" + k] = p.grad.detach().cpu().numpy()
    ms = None
    if time_it:
        for p in enc.params():
            p.grad = None
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        g = enc(a_t, A_t)
        at = enc.get_atom_array()
        loss = (at * w).sum() + g.sum()
        e1.record()
        loss.backward()
        e2.record()
        torch.cuda.synchronize()
        ms = (e0.elapsed_time(e1), e1.elapsed_time(e2))
    return out, ms


def cmp(a, b):
    worst = 0.0
    for k in a:
        d = np.abs(a[k] - b[k]).max() / max(np.abs(b[k]).max(), 1e-30)
        worst = max(worst, d)
        if d > 1e-4:
            print("   !!", k, d)
    return worst


if __name__ == "__main__":
    if "--waits" in sys.argv:
        import ctypes as C
        lib = gcnbmp._capi.lib
        run(128, 6, 64, 4144, True, True)            # warm
        dbg = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
        lib.bmp_debug_set_buffer_x3.argtypes = [C.c_void_p]
        lib.bmp_debug_set_buffer_x3(C.c_void_p(dbg.data_ptr()))
        run(128, 6, 64, 4144, True, True)
        lib.bmp_debug_set_buffer_x3(C.c_void_p(0))
        torch.cuda.synchronize()
        d = dbg.cpu().numpy().reshape(-1, 8)
        names = ["total", "mma:FULL", "mma:ACCE", "conv:EMPTY", "conv:work", "epi:ACCF", "epi:work", "mma:WFULL"]
        for i in list(range(3, 9)) + list(range(18, 24)):
            print("launch %2d: " % i + "  ".join("%s %d" % (n, v) for n, v in zip(names, d[i])))
        sys.exit(0)
    if "--prof" in sys.argv:
        run(128, 6, 64, 4144, True, True, time_it=True)
        sys.exit(0)
    for (H, T, N, mb, tied, nmax) in [(128, 6, 64, 300, True, None), (128, 3, 37, 130, False, 30), (64, 4, 64, 129, True, None),
                                      (256, 2, 50, 70, True, None)]:
        ref, _ = run(H, T, N, mb, tied, False, nmax)
        got, _ = run(H, T, N, mb, tied, True, nmax)
        print("H%d T%d N%d mb%d tied=%s: max rel diff x3 vs FFMA %.2e" % (H, T, N, mb, tied, cmp(got, ref)), flush=True)
    if "--time256" in sys.argv:
        for x3 in (False, True):
            _, ms = run(256, 8, 64, 2048, True, x3, time_it=True)
            print("H256 T8 N64 2048 molecules x3=%s: fwd %.2f ms bwd(+wgrad) %.2f ms" % (x3, ms[0], ms[1]), flush=True)
    if "--time" in sys.argv:
        for x3 in (False, True):
            _, ms = run(128, 6, 64, 4144, True, x3, time_it=True)
            print("H128 T6 N64 4144 molecules x3=%s: fwd %.2f ms bwd(+wgrad) %.2f ms" % (x3, ms[0], ms[1]), flush=True)
