"""Profiling target: BF16-mode encoder fwd+bwd (+readout) on 2048 molecules, H=128, T=6, N=64."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic

H, T, N, mb = 128, 6, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
atoms, adj = synthetic.random_molecules(rng, mb, N)
A, X = torch.tensor(adj).cuda(), torch.tensor(atoms).cuda()
net = gcnbmp.GGNNMono(H, H, T)
net.mode = gcnbmp.MODE_BF16
for i in range(3):
    net.cleargrads()
    g = net(X, A)
    (g.sum() + net.get_atom_array().sum()).backward()
torch.cuda.synchronize()
print("ok")
with torch.no_grad():
    for i in range(2):
        net(X, A)      # stash-free inference launch (ggnn_tc_kernel<128, false>)
torch.cuda.synchronize()
