#!/usr/bin/env python
"""bench.py -- drug pairs/s (fwd+bwd, GGNN + co-attention) on N B200s (BASELINE.json metric).

Workload (BASELINE.json configs[2]/[4], SURVEY.md 8d "C/E"): GGNN hidden 128, T = 6 tied,
4 bond types, gated readout O = 128, Nie fine-grained co-attention (head 8, tanh), HolE head
-> 86 logits, sigmoid cross-entropy; one STEP = forward + backward + (N>1: one NCCL allreduce
of the flat gradient) + Adam over a GLOBAL batch of 65 536 synthetic pairs padded to 64 atoms
(strong scaling: each of N ranks owns 65 536 / N pairs).

  value  : pairs/s with this rank's inputs already resident in HBM (CUDA events, max over ranks)
  e2e    : the same step through the public API with HOST (pinned) inputs, H2D inside the timed
           region (streamed per micro-batch on a copy stream) and the loss read back (D2H)
  roofline / cpu_baseline : see DESIGN.md "Measurement".

`--impl reference` times the NumPy oracle (the restated Chainer CPU path; Chainer itself is not
installable here) on a bounded sample of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gcn-bmp_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CFG = dict(H=128, T=6, N=64, E=4, O=128, head=8, K=86)
GLOBAL_BATCH = 65536
METRIC = "drug pairs/sec (fwd+bwd, GGNN+co-attention)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        threading.Thread.__init__(self, daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def build_model(mode="bf16"):
    import gcnbmp
    gcnbmp.seed(777)
    enc = gcnbmp.GGNNMono(CFG["O"], CFG["H"], CFG["T"], weight_tying=True)
    enc.mode = gcnbmp.MODE_BF16 if mode == "bf16" else gcnbmp.MODE_F32
    BF = enc.mode
    attn = gcnbmp.NieFineCoattention(CFG["H"], CFG["O"], CFG["head"], activation=gcnbmp.functions.tanh)
    mlp = gcnbmp.HolE(CFG["K"], hidden_dims=())
    mlp.l_out.ensure(CFG["O"])
    attn.mode = BF      # BF16 mode: the co-attention's (H,H)/(O,H) weight-gradient contractions run on tcgen05
    return gcnbmp.GraphConvPredictorForPair(enc, attn, mlp)


def host_batch(n_pairs, seed, unique=4096):
    """n_pairs synthetic pairs in pinned host memory (a pool of `unique` generated pairs tiled:
    timing does not depend on content, generation time does)."""
    import torch
    from gcnbmp import synthetic
    u = min(unique, n_pairs)
    a1, A1, a2, A2, y = synthetic.random_pairs(seed, u, CFG["N"], CFG["K"])
    reps = (n_pairs + u - 1) // u
    out = []
    for arr in (a1, A1, a2, A2, y):
        t = torch.empty((n_pairs,) + arr.shape[1:], dtype=torch.from_numpy(arr).dtype).pin_memory()
        src = torch.from_numpy(arr)
        for r in range(reps):
            s, e = r * u, min(n_pairs, (r + 1) * u)
            t[s:e].copy_(src[: e - s])
        out.append(t)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import gcnbmp
    from gcnbmp.train import PairTrainer, algorithmic_flops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    gcnbmp._capi.check(gcnbmp._capi.lib.bmp_device_check())
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_local = args.pairs // world
    model = build_model(args.mode)
    trainer = PairTrainer(model, chunk=args.chunk, world_size=world, alpha=1e-3)
    if world > 1:   # replicas start from rank 0's parameters
        dist.broadcast(trainer.flat, src=0)
    host = host_batch(n_local, seed=2018 + rank)
    resident = [t.to(dev) for t in host]
    gcount = float(args.pairs * CFG["K"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        gcnbmp.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, gcnbmp.launch_count()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- value: inputs resident in HBM ----
    ms, launches = timed(lambda: trainer.step(*resident, global_count=gcount), args.steps, args.warmup)
    ms_per_step = ms / args.steps
    value = args.pairs / (ms_per_step * 1e-3)
    # ---- e2e: host inputs, H2D inside the timed region, loss read back ----
    losses = []

    def e2e_step():
        trainer.h2d_bytes = 0
        loss = trainer.step(*host, global_count=gcount)
        losses.append(float(loss.item()))       # D2H of the step's result

    e2e_steps = max(1, args.e2e_steps if args.e2e_steps else min(args.steps, 3))
    e2e_ms, _ = timed(e2e_step, e2e_steps, max(1, min(args.warmup, 3)))
    e2e_value = args.pairs / (e2e_ms / e2e_steps * 1e-3)
    sampler.stop_flag = True
    h2d = trainer.h2d_bytes * world
    # ---- informational: the same end-to-end step with the 0/1 adjacency held as uint8 in pinned memory (4x fewer PCIe bytes;
    # the tcgen05 kernels stage bytes directly).  `e2e` above keeps the reference's fp32 adjacency and is the headline. ----
    e2e_u8 = None
    if args.e2e_u8:
        host_u8 = [t if i not in (1, 3) else t.to(torch.uint8).pin_memory() for i, t in enumerate(host)]

        def e2e_u8_step():
            trainer.h2d_bytes = 0
            losses.append(float(trainer.step(*host_u8, global_count=gcount).item()))

        u8_ms, _ = timed(e2e_u8_step, e2e_steps, 1)
        e2e_u8 = dict(value=round(args.pairs / (u8_ms / e2e_steps * 1e-3), 1), unit="pairs/s",
                      h2d_bytes_per_step=int(trainer.h2d_bytes * world),
                      note="host adjacency held as uint8 (exact for 0/1 bonds), staged by the tcgen05 kernels as is; not the headline")
        del host_u8

    # ---- roofline of the dominant kernel (fused GGNN encoder forward), timed live ----
    fl = algorithmic_flops(CFG["H"], CFG["T"], CFG["N"], CFG["E"], CFG["O"], CFG["head"], CFG["K"])
    roof = None
    if rank == 0:
        nmol = min(n_local, args.chunk)
        enc = model.graph_conv
        a1, A1 = resident[0][:nmol], resident[1][:nmol]
        with torch.no_grad():
            for _ in range(2):
                enc(a1, A1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                enc(a1, A1)
            e1.record()
            torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps      # encoder launch + (small) readout launch
        pk = peaks()
        achieved = nmol * fl["encoder"] / (k_ms * 1e-3) / 1e12
        roof = dict(bound="tensor", kernel=("ggnn_tc_kernel<128>" if args.mode == "bf16" else "ggnn_fwd_kernel<2>") + " (+readout launch)",
                    achieved=round(achieved, 3), peak=pk["bf16_sustained"], unit="TFLOP/s",
                    frac=round(achieved / pk["bf16_sustained"], 5), traffic=(231.1e6 if args.mode == "bf16" else 229.0e6) * nmol / 2048.0,
                    peak_source=pk["source"] + " bf16 sustained (kernel timed inside a long step)",
                    note="algorithmic FLOPs = 188.8 MFLOP x %d molecules per launch; traffic = ncu dram bytes r+w per 2048-molecule launch, scaled to this launch" % nmol)
    # ---- the parity-exact fp32 mode on the same workload (1 warm-up + 1 step), for the record ----
    fp32_exact = None
    if args.mode == "bf16" and not args.no_fp32:
        import gcnbmp as _g
        model.graph_conv.mode = model.attn.mode = _g.MODE_F32
        ms32, _ = timed(lambda: trainer.step(*resident, global_count=gcount), 1, 1)
        model.graph_conv.mode = model.attn.mode = _g.MODE_BF16
        fp32_exact = dict(value=round(args.pairs / (ms32 * 1e-3), 1), unit="pairs/s", ms_per_step=round(ms32, 3),
                          note="BMP_MODE_F32: parity <= 1e-4 vs the oracle")
    # ---- informational: BASELINE config D (GGNN H256 T8 + R1 readout + HolE->1, forward only) on this rank's GPU, same inputs ----
    config_d = None
    if rank == 0 and args.mode == "bf16" and not args.no_config_d:
        nd = min(n_local, 4096)
        encD = gcnbmp.GGNN(256, hidden_dim=256, n_layers=8, weight_tying=True)
        mD = gcnbmp.GraphConvPredictorForPair(encD, None, gcnbmp.HolE(1, hidden_dims=()))
        encD.mode = gcnbmp.MODE_BF16
        argsD = [t[:nd] for t in resident[:4]]
        with torch.no_grad():
            for _ in range(3):
                mD(*argsD)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                mD(*argsD)
            e1.record()
            torch.cuda.synchronize()
        d_ms = e0.elapsed_time(e1) / 5
        fd = algorithmic_flops(256, 8, CFG["N"], CFG["E"], 256, K=1, readout="r1", attn=False, D=256)
        config_d = dict(value=round(nd / (d_ms * 1e-3), 1), unit="pairs/s per GPU", ms=round(d_ms, 3), pairs=nd,
                        achieved_tflops=round(nd * fd["pair_fwd"] / (d_ms * 1e-3) / 1e12, 1),
                        frac_of_bf16_sustained_peak=round(nd * fd["pair_fwd"] / (d_ms * 1e-3) / 1e12 / peaks()["bf16_sustained"], 4),
                        note="forward only, inputs resident, hidden-256 tcgen05 encoder + readout (csrc/ggnn_tc256.cu); not the headline")
        del mD, encD
    cpu = cpu_baseline(args) if rank == 0 and not args.no_cpu else None
    if rank == 0:
        line = dict(metric=METRIC, value=round(value, 1), unit="pairs/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=round(ms_per_step, 3), higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="bf16" if args.mode == "bf16" else "f32", data="synthetic",
                    config=dict(workload="GGNN(H128,T6,tied,E4,N64 padded)+Nie co-attention(head8,tanh,O128)+HolE->86, "
                                         "sigmoid-CE, fwd+bwd+Adam, global batch %d pairs" % args.pairs,
                                global_batch=args.pairs, micro_batch=args.chunk, parallelism="dp%d" % world,
                                l2="inputs (%.1f GB/rank) larger than L2" % (sum(t.numel() * t.element_size() for t in resident) / 1e9),
                                mode=("bf16 operands on tcgen05 for the GGNN encoder fwd/bwd/wgrad, the gated readout and the co-attention "
                                      "(fp32 accumulate, state, softmaxes), HolE/head/loss fp32; atom states <= 5e-2 max-rel / 1e-2 rms-rel vs the fp64 oracle"
                                      if args.mode == "bf16" else "fp32-exact (parity <= 1e-4 vs oracle)")),
                    e2e=dict(value=round(e2e_value, 1), unit="pairs/s", h2d_bytes_per_step=int(h2d),
                             d2h_bytes_per_step=4 * world, loss=losses[-1] if losses else None),
                    gpu_launches=int(launches), clocks=sampler.summary(), roofline=roof, cpu_baseline=cpu,
                    fp32_exact=fp32_exact, e2e_u8=e2e_u8, config_d_forward=config_d,
                    flops_per_pair_fwd=fl["pair_fwd"], achieved_tflops_step=round(3 * fl["pair_fwd"] * value / 1e12, 3))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def oracle_step_fn(n_pairs, seed=2018):
    """The oracle (NumPy restatement of the Chainer op sequence) fwd+bwd+Adam on n_pairs of the
    bench workload, fp32, BLAS on all host cores."""
    import cases
    from oracle import reference_path as R
    from gcnbmp import synthetic
    spec = dict(enc="mono", H=CFG["H"], T=CFG["T"], tied=True, sum_readout=False, O=CFG["O"], attn="nie",
                head=CFG["head"], hole_hidden=(), K=CFG["K"])
    shapes = {"graph_conv/" + k: v for k, v in R.ggnn_mono_shapes(CFG["O"], CFG["H"], CFG["T"]).items()}
    shapes.update({"attn/" + k: v for k, v in R.coattn_shapes(CFG["H"], CFG["O"], CFG["head"]).items()})
    shapes.update({"mlp/" + k: v for k, v in R.hole_shapes(CFG["O"], CFG["K"], ()).items()})
    params = R.init_params(shapes, np.random.default_rng(777), dtype=np.float32)
    table = R.wrap_params(params, dtype=np.float32)
    model = cases.oracle_model(spec, table)
    a1, A1, a2, A2, y = synthetic.random_pairs(seed, n_pairs, CFG["N"], CFG["K"])
    m = {k: np.zeros_like(v.data) for k, v in table.items()}
    v2 = {k: np.zeros_like(v.data) for k, v in table.items()}
    state = dict(t=0)

    def step():
        loss, _, grads = R.loss_and_grads(model, table, (a1, A1, a2, A2), y)
        state["t"] += 1
        t = state["t"]
        a_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        for k, p in table.items():
            g = grads[k]
            m[k] += 0.1 * (g - m[k])
            v2[k] += 0.001 * (g * g - v2[k])
            p.data -= (a_t * m[k] / (np.sqrt(v2[k]) + 1e-8)).astype(np.float32)
        return float(loss)
    return step


def cpu_baseline(args, sample=64, iters=2):
    step = oracle_step_fn(sample)
    step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = (time.perf_counter() - t0) / iters
    return dict(value=round(sample / dt, 2), unit="pairs/s", cores=os.cpu_count(), kind="port",
                sample="%d pairs of the bench workload x %d iterations, fp32 NumPy oracle (BLAS threads = all cores)" % (sample, iters))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 128
    step = oracle_step_fn(sample)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = sample / dt
    desc = "%d pairs per step (bounded sample of the 65536-pair workload), fp32 NumPy oracle, BLAS on all cores" % sample
    line = dict(impl="reference", metric=METRIC, value=round(value, 2), unit="pairs/s",
                n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=steps, warmup=min(args.warmup, 1),
                ms_per_step=round(dt * 1e3, 2), higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload="GGNN(H128,T6,tied,E4,N64 padded)+Nie co-attention(head8,tanh,O128)+HolE->86, "
                                     "sigmoid-CE, fwd+bwd+Adam", sample_pairs=sample,
                            note="Chainer is not installable here (Python 2 reference, no network): the NumPy oracle "
                                 "executes the reference's op sequence"),
                cpu_baseline=dict(value=round(value, 2), unit="pairs/s", cores=os.cpu_count(), kind="port", sample=desc),
                e2e=dict(value=round(value, 2), unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=GLOBAL_BATCH, help="global batch (pairs per step)")
    ap.add_argument("--chunk", type=int, default=4144,
                    help="micro-batch (pairs) per forward/backward; 4144 molecules = 2072 two-molecule tiles = 14 full waves of 148 CTAs")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (0: min(--steps, 3))")
    ap.add_argument("--no-e2e-u8", dest="e2e_u8", action="store_false", help="skip the informational uint8-adjacency e2e measurement")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config-d", action="store_true", help="skip the informational config-D forward measurement")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: GGNN encoder fwd/bwd/wgrad on tcgen05 (stated bound); fp32: parity <= 1e-4 path")
    ap.add_argument("--no-fp32", action="store_true", help="skip the extra fp32-exact measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
