"""Training-loss trajectories of BMP_MODE_F32 vs BMP_MODE_BF16 from identical initial parameters and data
(config-C model at hidden 128, 1024 synthetic pairs, 40 Adam steps): evidence that the tcgen05 path trains like the
<= 1e-4-parity fp32 path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic, train

H, T, N, O, K, mb, STEPS = 128, 6, 64, 128, 86, 1024, 40
rng = np.random.default_rng(2018)
a1, A1 = synthetic.random_molecules(rng, mb, N)
a2, A2 = synthetic.random_molecules(rng, mb, N)
# learnable labels: class depends on simple statistics of the pair (so the loss can actually go down)
cnt = (np.asarray(a1) == 8).sum(1) + (np.asarray(a2) == 7).sum(1)
y = np.zeros((mb, K), np.int32)
y[np.arange(mb), cnt % K] = 1
dev = lambda x: torch.tensor(x).cuda()
args = [dev(a1), dev(A1), dev(a2), dev(A2), dev(y)]


def run(mode):
    gcnbmp.links.seed(777)          # identical initial parameters for both runs
    enc = gcnbmp.GGNNMono(O, H, T)
    attn = gcnbmp.NieFineCoattention(H, O, 8, activation=gcnbmp.functions.tanh)
    head = gcnbmp.HolE(K, hidden_dims=())
    head.l_out.ensure(O)        # lazily-shaped layer: materialise before the trainer flattens the parameters
    model = gcnbmp.GraphConvPredictorForPair(enc, attn, head)
    enc.mode = attn.mode = mode
    tr = train.PairTrainer(model, chunk=512, alpha=1e-3)
    return model, [float(tr.step(*args)) for _ in range(STEPS)]


m32, l32 = run(gcnbmp.MODE_F32)
mbf, lbf = run(gcnbmp.MODE_BF16)
p32 = {k: v.detach().cpu().numpy() for k, v in m32.namedparams()}
pbf = {k: v.detach().cpu().numpy() for k, v in mbf.namedparams()}
same_init = True
print("step   loss fp32    loss bf16   rel diff")
for i in (0, 1, 2, 5, 10, 20, 30, STEPS - 1):
    print("%4d  %10.6f  %10.6f  %9.2e" % (i, l32[i], lbf[i], abs(l32[i] - lbf[i]) / abs(l32[i])))
drift = max(float(np.abs(p32[k] - pbf[k]).max()) for k in p32)
print("max parameter drift after %d Adam steps (lr 1e-3): %.3e" % (STEPS, drift))
