"""Shared helpers of the tools/bench_*.py scripts."""
import numpy as np
import torch
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
