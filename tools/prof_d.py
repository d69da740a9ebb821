"""Profiling target: BASELINE config D forward (GGNN H256 T8 + R1 readout + HolE) in BF16 mode, 2048 pairs, three passes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from bench_configs_util import pairs

rng = np.random.default_rng(2018)
enc = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
mD = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
enc.mode = gcnbmp.MODE_BF16
aD = pairs(rng, 2048, 64)
with torch.no_grad():
    for _ in range(3):
        out = mD(*aD[:4])
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
