"""Profiling target: one pair training micro-batch sequence (config C shapes) through PairTrainer; BF16 mode, or MODE_F32
with BMP_PROF_FP32=1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic, train

H, T, N, O, K, mb = 128, 6, 64, 128, 86, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
a1, A1 = synthetic.random_molecules(rng, mb, N)
a2, A2 = synthetic.random_molecules(rng, mb, N)
y = (rng.random((mb, K)) < 0.1).astype(np.int32)
enc = gcnbmp.GGNNMono(O, H, T)
attn = gcnbmp.NieFineCoattention(H, O, 8, activation=gcnbmp.functions.tanh)
head = gcnbmp.HolE(K, hidden_dims=())
head.l_out.ensure(O)        # lazily-shaped layer: materialise before the trainer flattens the parameters
model = gcnbmp.GraphConvPredictorForPair(enc, attn, head)
enc.mode = attn.mode = gcnbmp.MODE_F32 if os.environ.get("BMP_PROF_FP32") else gcnbmp.MODE_BF16
tr = train.PairTrainer(model, chunk=min(mb, 4144))
dev = lambda x: torch.tensor(x).cuda()
args = [dev(a1), dev(A1), dev(a2), dev(A2), dev(y)]
for i in range(3):
    loss = tr.step(*args)
torch.cuda.synchronize()
print("ok", float(loss))
