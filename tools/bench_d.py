"""BASELINE config D on one B200: GGNN H256 T8 (tied) + R1 readout O=256 + HolE->1, forward only, inputs resident.
fp32 kernels vs the hidden-256 tcgen05 kernel (csrc/ggnn_tc256.cu); CUDA events, median of 5 after 3 warm-ups."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic
from bench_configs_util import timed, pairs

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(2018)
enc = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
mD = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
aD = pairs(rng, NP, 64)
aD8 = [aD[0], aD[1].to(torch.uint8), aD[2], aD[3].to(torch.uint8)]
F_PAIR = 1879.2e6       # algorithmic FLOPs per pair, forward, config D


def fwd(args):
    def fn():
        with torch.no_grad():
            return mD(*args[:4])
    return fn


ref = fwd(aD)().clone()
ms = timed(fwd(aD))
print("D  fp32 kernels:            %8.3f ms  %10.0f pairs/s  %6.1f TFLOP/s" % (ms, NP / ms * 1e3, F_PAIR * NP / ms / 1e9))
enc.mode = gcnbmp.MODE_BF16
got = fwd(aD)().clone()
err = float((got - ref).abs().max() / ref.abs().max())
ms = timed(fwd(aD))
print("D  BF16 (tcgen05, H256):    %8.3f ms  %10.0f pairs/s  %6.1f TFLOP/s   logits max-rel vs fp32 %.2e" % (ms, NP / ms * 1e3, F_PAIR * NP / ms / 1e9, err))
ms = timed(fwd(aD8))
print("D  BF16, uint8 adjacency:   %8.3f ms  %10.0f pairs/s  %6.1f TFLOP/s" % (ms, NP / ms * 1e3, F_PAIR * NP / ms / 1e9))
