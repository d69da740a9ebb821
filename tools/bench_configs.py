"""Throughput of the OTHER BASELINE.json configurations (parity-test cases, not bench.py lines) on one B200:
A  GGNN H32 T4 tied, sum readout, HolE->1, 128 pairs N<=50, fwd+bwd           (fp32 kernels; H=32 is below the tcgen05 tiles)
B  RelGCN 64->64 x4, scale_adj, readout O=64, HolE->1, 4096 pairs N<=64, fwd+bwd (fp32 kernels and tcgen05)
D  GGNN H256 T8 + R1 readout O=256 + HolE->1, forward only                     (fp32 kernels and the hidden-256 tcgen05 kernels)
Inputs resident on the device, CUDA events, median of 5 after 3 warm-ups."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import gcnbmp
from gcnbmp import synthetic


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def pairs(rng, mb, N):
    a1, A1 = synthetic.random_molecules(rng, mb, N)
    a2, A2 = synthetic.random_molecules(rng, mb, N)
    y = (rng.random((mb, 1)) < 0.33).astype(np.int32)
    return [torch.tensor(x).cuda() for x in (a1, A1, a2, A2, y)]


def train_step(model, args):
    def fn():
        model.cleargrads()
        loss = gcnbmp.sigmoid_cross_entropy(model(*args[:4]), args[4])
        loss.backward()
    return fn


rng = np.random.default_rng(2018)
f = gcnbmp.functions
# A
enc = gcnbmp.GGNNMono(32, 32, 4, weight_tying=True, sum_readout=True)
mA = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
aA = pairs(rng, 128, 50)
ms = timed(train_step(mA, aA))
print("A  GGNN H32 T4 sum-readout, 128 pairs fwd+bwd (fp32):   %8.3f ms  %10.0f pairs/s" % (ms, 128 / ms * 1e3))
from gcnbmp import train as _train
for nb in (128, 32):
    sub = [t[:nb] for t in aA]
    for graph in (False, True):
        tr = _train.PairTrainer(mA, chunk=nb, optimizer=True, graph=graph, alpha=1e-3)
        ms = timed(lambda: tr.step(*sub), reps=20)
        print("A  PairTrainer step (fwd+bwd+Adam), %3d pairs, %s: %8.3f ms  %10.0f pairs/s" % (nb, "CUDA graph" if graph else "eager     ", ms, nb / ms * 1e3))
# B
enc = gcnbmp.RelGCN(64, ch_list=[64, 64, 64, 64, 64], scale_adj=True)
mB = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
aB = pairs(rng, 4096, 64)
ms = timed(train_step(mB, aB))
print("B  RelGCN 64x4, 4096 pairs fwd+bwd (fp32):              %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
enc.mode = gcnbmp.MODE_BF16
ms = timed(train_step(mB, aB))
print("B  RelGCN 64x4, 4096 pairs fwd+bwd (BF16, tcgen05):     %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
from gcnbmp import functional as Fn
gflat = mB.flatten_parameters()[1]


def sink_step():
    gflat.zero_()
    loss = gcnbmp.sigmoid_cross_entropy(mB(*aB[:4]), aB[4])
    Fn.set_grad_sink(True)
    Fn.set_weight_cache(True)
    try:
        loss.backward()
    finally:
        Fn.set_grad_sink(False)


ms = timed(sink_step)
Fn.set_weight_cache(False)
print("B  same, gradient sink + cached weight images:          %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
# D
enc = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
mD = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
aD = pairs(rng, 4096, 64)


def fwd():
    with torch.no_grad():
        mD(*aD[:4])


ms = timed(fwd)
print("D  GGNN H256 T8 + R1 + HolE, 4096 pairs forward (fp32): %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
enc.mode = gcnbmp.MODE_BF16
ms = timed(fwd)
print("D  GGNN H256 T8 + R1 + HolE, 4096 pairs forward (BF16, tcgen05): %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
# C (inference): the bench workload forward only, BF16 mode -- the shape class of config D at the hidden size the tcgen05 kernels cover
enc = gcnbmp.GGNNMono(128, 128, 6)
attn = gcnbmp.NieFineCoattention(128, 128, 8, activation=f.tanh)
mC = gcnbmp.GraphConvPredictorForPair(enc, attn, gcnbmp.HolE(86, hidden_dims=()))
enc.mode = attn.mode = gcnbmp.MODE_BF16
aC = pairs(rng, 4096, 64)


def fwdC():
    with torch.no_grad():
        mC(*aC[:4])


ms = timed(fwdC)
print("C  GGNN H128 T6 + co-attention + HolE->86, 4096 pairs FORWARD (BF16): %8.3f ms  %10.0f pairs/s" % (ms, 4096 / ms * 1e3))
# f-1: pairs as index pairs into a device-resident drug table (KAIST-sized: 1704 unique drugs), 65 536 pairs per step, BF16 mode
U, NP = 1704, 65536
tab_a, tab_A = synthetic.random_molecules(rng, U, 64)
tab_a, tab_A = torch.tensor(tab_a).cuda(), torch.tensor(tab_A.astype(np.uint8)).cuda()
i1, i2 = rng.integers(0, U, NP), rng.integers(0, U, NP)
yK = (rng.random((NP, 86)) < 0.05).astype(np.int32)
head = gcnbmp.HolE(86, hidden_dims=())
head.l_out.ensure(128)
encI = gcnbmp.GGNNMono(128, 128, 6)
attnI = gcnbmp.NieFineCoattention(128, 128, 8, activation=f.tanh)
mI = gcnbmp.GraphConvPredictorForPair(encI, attnI, head)
encI.mode = attnI.mode = gcnbmp.MODE_BF16
trI = _train.PairTrainer(mI, chunk=2048, alpha=1e-3)
for dd in (False, True):
    ms = timed(lambda: trI.step_indexed(tab_a, tab_A, i1, i2, yK, dedupe=dd), reps=3, warm=2)
    print("E  index pairs over a %d-drug device table, %d pairs/step, fwd+bwd+Adam, %s: %8.2f ms  %10.0f pairs/s  (H2D %.2f MB/step)"
          % (U, NP, "each drug encoded once" if dd else "per-pair encoding     ", ms, NP / ms * 1e3, trI.h2d_bytes / 1e6))
# D via the drug table: GGNN H256 T8 + R1 readout + HolE, 1 M index pairs over 1704 drugs, forward only (fp32 kernels, then BF16 mode)
tab_Af = tab_A.float()
encD = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
headD = gcnbmp.HolE(1, hidden_dims=())
headD.l_out.ensure(256)
trD = _train.PairTrainer(gcnbmp.GraphConvPredictorForPair(encD, None, headD), chunk=65536, optimizer=False)
j1, j2 = rng.integers(0, U, 1 << 20), rng.integers(0, U, 1 << 20)
ms = timed(lambda: trD.predict_indexed(tab_a, tab_Af, j1, j2), reps=3, warm=1)
print("D  GGNN H256 T8 + R1 + HolE, 1 M index pairs over a %d-drug table, forward (fp32): %8.2f ms  %10.0f pairs/s" % (U, ms, (1 << 20) / ms * 1e3))
encD.mode = gcnbmp.MODE_BF16
ms = timed(lambda: trD.predict_indexed(tab_a, tab_A, j1, j2), reps=3, warm=1)
print("D  same, encoder + readout on the hidden-256 tcgen05 kernels (byte adjacency):            %8.2f ms  %10.0f pairs/s" % (ms, (1 << 20) / ms * 1e3))
