#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]): multipassGAN-4x generator + spatial discriminator,
16x16 -> 64x64 tile batches, one loop body = discRuns D steps + genRuns G steps, data parallel (one NCCL
all-reduce of the flat gradient per optimizer step).

  python tools/bench_train.py [--batch 16] [--steps 10] [--warmup 3]          # 1 GPU
  python -m torch.distributed.run --nproc-per-node N tools/bench_train.py    # N GPUs (weak scaling: batch per rank)
Prints one JSON line: iterations/s and tiles/s (aggregate), algorithmic TFLOP/s, and the CPU oracle
(torch-CPU fp32 autograd) on the same loop body as `cpu_baseline`.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

import mpgan_b200  # noqa: F401
from mpgan_b200 import parallel as par, training as T

# per sample G fwd 2.60 GMAC, D fwd 0.051 GMAC; D step = G fwd + 2 D fwd + 2 D bwd(2x);
# G step = G fwd + 2 D fwd + D bwd (dgrad only ~1x) + G bwd (2x)
G_FWD, D_FWD = 2 * 2.60e9 * (4096 / 4096), 2 * 0.0514e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--disc-runs", type=int, default=1)
    ap.add_argument("--gen-runs", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp16", choices=["fp32", "fp16"])
    ap.add_argument("--graphs", type=int, default=1, help="replay each optimizer step as a captured CUDA graph (single GPU)")
    args = ap.parse_args()
    rank, local, world = par.init_from_env()
    torch.cuda.set_device(local)
    L, u, B = 16, 4, args.batch
    S = L * u
    tr = T.Trainer4x(L, u, B, seed=1, device=local, precision=args.precision, graphs=bool(args.graphs))
    rng = np.random.default_rng(100 + rank)
    xs = torch.from_numpy(rng.random((B, L * L * 4), dtype=np.float32)).pin_memory()
    ys = torch.from_numpy(rng.random((B, S * S), dtype=np.float32)).pin_memory()

    def body():
        xd, yd = xs.cuda(non_blocking=True), ys.cuda(non_blocking=True)  # H2D of the tile batch every loop body
        return tr.iteration([(xd, yd)] * args.disc_runs, [(xd, yd)] * args.gen_runs)  # .item() of the losses = D2H

    for _ in range(args.warmup):
        body()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = tr.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = body()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    it_ms = ms / args.steps
    flop_it = B * (args.disc_runs * (G_FWD + 2 * D_FWD * 3) + args.gen_runs * (G_FWD * 3 + 2 * D_FWD + 2 * D_FWD))
    line = dict(metric="training loop bodies/sec (4x G + spatial D, tiles 16x16->64x64)", value=world * 1e3 / it_ms,
                unit="iteration/s", tiles_per_s=world * B * 1e3 / it_ms, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=it_ms, higher_is_better=True, scaling="weak",
                dtype="f32" if args.precision == "fp32" else "f16 fwd / bf16 dgrad + wgrad on tcgen05, f32 thin layers + optimizer", data="synthetic",
                config=dict(workload="multipassGAN-4x training step (BASELINE.json configs[3])", batch_per_gpu=B,
                            discRuns=args.disc_runs, genRuns=args.gen_runs, tile="16x16 -> 64x64", cuda_graphs=bool(tr.use_graphs)),
                algorithmic_tflops=world * flop_it / (it_ms * 1e-3) / 1e12,
                gpu_launches=int((tr.launches - l0)), losses=out,
                e2e=dict(value=world * 1e3 / it_ms, unit="iteration/s", h2d_bytes_per_step=int(xs.numel() * 4 + ys.numel() * 4),
                         d2h_bytes_per_step=64 * (args.disc_runs + args.gen_runs)))
    if not args.no_cpu_baseline:
        from oracle import networks as on, training as ot
        cfg = on.make_cfg_4x(L, upRes=u, upsampling_mode=2, batch_norm=True)
        values = {}
        hp = dict(kk=5.0, kk2=1e-5, seed=1)
        od, og_ = ot.Adam(2e-4, 0.5), ot.Adam(2e-4, 0.5)
        xb, yb = xs.numpy(), ys.numpy()
        ot.train_iteration(values, [(xb, yb)], [(xb, yb)], cfg, hp, od, og_, dtype=torch.float32)  # warm-up / var creation
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            ot.train_iteration(values, [(xb, yb)] * args.disc_runs, [(xb, yb)] * args.gen_runs, cfg, hp, od, og_,
                               dtype=torch.float32)
        dt = (time.perf_counter() - t0) / n
        line["cpu_baseline"] = dict(value=1.0 / dt, unit="iteration/s", cores=torch.get_num_threads(), kind="port",
                                    sample="oracle port (torch-CPU fp32 autograd) of the same loop body, %d iterations" % n)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
