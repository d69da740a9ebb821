"""BASELINE config D: GGNN H256 T8 (tied) + R1 readout O=256 + HolE->1, 1 M synthetic pairs, forward only, sharded over N GPUs.
No data-path collective: each rank scores its contiguous share of the pairs; one barrier on both sides of the timed
region, CUDA events, max over ranks.  Inputs: a device-resident pool of 16 384 synthetic pairs per rank in the reference layout
(atoms int32, adj fp32 (mb,4,64,64) = 2.1 GB, larger than L2), cycled in micro-batches of 4096 pairs.
  python tools/bench_d_sharded.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_d_sharded.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
import gcnbmp
from bench_configs_util import pairs

TOTAL = int(os.environ.get("D_PAIRS", 1 << 20))
POOL, CHUNK = 16384, 4096
F_PAIR = 1879.2e6       # algorithmic FLOPs per pair, forward, config D

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
rng = np.random.default_rng(2018 + rank)
enc = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
model = gcnbmp.GraphConvPredictorForPair(enc, None, gcnbmp.HolE(1, hidden_dims=()))
enc.mode = gcnbmp.MODE_BF16
pool = pairs(rng, POOL, 64)
share = TOTAL // world
out = torch.empty((share, 1), device="cuda")


def run(n):
    with torch.no_grad():
        for lo in range(0, n, CHUNK):
            m = min(CHUNK, n - lo)
            o = lo % POOL
            out[lo:lo + m] = model(pool[0][o:o + m], pool[1][o:o + m], pool[2][o:o + m], pool[3][o:o + m])


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


run(3 * CHUNK)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(share)
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms)
if rank == 0:
    v = share * world / ms * 1e3
    print(json.dumps({"metric": "drug pairs/sec (forward, GGNN H256 T8 + R1 readout + HolE)", "value": round(v, 1), "unit": "pairs/s",
                      "n_gpus": world, "ms": round(ms, 2), "pairs": share * world, "scaling": "strong", "dtype": "bf16",
                      "algorithmic_tflops": round(v * F_PAIR / 1e12, 1),
                      "frac_of_bf16_peak_per_gpu": round(v * F_PAIR / 1e12 / world / 1658.7, 3),
                      "config": {"workload": "BASELINE config D, per-pair encoding, inputs resident (16384-pair pool per rank, 2.1 GB > L2)"}}))
if world > 1:
    dist.destroy_process_group()
