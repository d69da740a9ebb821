"""Debug helper: BF16-mode (tcgen05 fwd + bwd + wgrad) gradients vs the fp32 kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic

H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 5
N = int(sys.argv[4]) if len(sys.argv) > 4 else 64
tied = (sys.argv[5] != "untied") if len(sys.argv) > 5 else True
rng = np.random.default_rng(0)
atoms, adj = synthetic.random_molecules(rng, mb, N)
net = gcnbmp.GGNNMono(H, H, T, weight_tying=tied)
w1 = torch.tensor(rng.standard_normal((mb, H)), dtype=torch.float32, device="cuda")
w2 = torch.tensor(rng.standard_normal((mb, N, H)), dtype=torch.float32, device="cuda")
res = {}
for mode in (gcnbmp.MODE_F32, gcnbmp.MODE_BF16):
    net.mode = mode
    net.cleargrads()
    g = net(atoms, adj)
    loss = (g * w1).sum() + (net.get_atom_array() * w2).sum()
    loss.backward()
    torch.cuda.synchronize()
    res[mode] = net.grad_dict()
for k in res[gcnbmp.MODE_F32]:
    a, b = res[gcnbmp.MODE_BF16][k], res[gcnbmp.MODE_F32][k]
    rms = np.sqrt(((a - b) ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-30)
    print("%-28s rms rel err %.3e   max|ref| %.3e  nan=%d" % (k, rms, np.abs(b).max(), int(np.isnan(a).sum())))
if len(sys.argv) > 6:
    mbig = int(sys.argv[6])
    atoms, adj = synthetic.random_molecules(rng, mbig, N)
    A, X = torch.tensor(adj).cuda(), torch.tensor(atoms).cuda()
    for mode, name in ((gcnbmp.MODE_F32, "fp32"), (gcnbmp.MODE_BF16, "tcgen05 bf16")):
        net.mode = mode
        for i in range(3):
            if i == 1:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            net.cleargrads()
            g = net(X, A)
            (g.sum() + net.get_atom_array().sum()).backward()
        e1.record()
        torch.cuda.synchronize()
        print("%s: encoder+readout fwd+bwd of %d molecules: %.2f ms" % (name, mbig, e0.elapsed_time(e1) / 2))
