"""Loop body of the 8x trainer at the shipped first-network configuration (GAN/example_run_training.py:4: tile 16 -> 128,
startFms / maxFms 256, 3x3 filters, batch 16, lambda_t 1.0): critic step + temporal-critic step + generator step, CUDA events.
python tools/bench_train8x.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import training8x as t8

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, L, S, C = 16, 16, 128, 6
tr = t8.Trainer8x(L, 8, C, 256, 256, 3, batch=B, learning_rate=1e-4, lambda_t=1.0)
dev = tr.cx.device
g = torch.Generator().manual_seed(1)
xs, ys = torch.rand((B, L * L * C), generator=g).to(dev), torch.rand((B, S * S), generator=g).to(dev)
nt = (B // 3) * 3
xt, yt = torch.rand((nt, L * L * C), generator=g).to(dev), torch.rand((nt, S * S), generator=g).to(dev)
lf, lft = torch.rand((B, 1)), torch.rand((nt // 3, 1))


def body(pct=3.0, z=2):
    tr.disc_step(xs, ys, pct, z, lf)
    tr.t_disc_step(xt, yt, pct, z, lft)
    tr.gen_step(xs, ys, pct, z, xt, yt)


body()
torch.cuda.synchronize()
l0 = tr.cx.launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    body()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
params = sum(p.total for p in (tr.gen.ps, tr.disc.ps, tr.tdisc.ps))
print(json.dumps({"metric": "8x trainer loop body (critic + temporal critic + generator step), fp32 CUDA-core kernels", "ms_per_body": ms,
                  "bodies_per_s": 1000.0 / ms, "batch": B, "tile": "16 -> 128", "fms": 256, "parameters": params,
                  "train_calls_per_body": (tr.cx.launches - l0) / steps, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
