"""Data-parallel step of the 8x trainer (torchrun, one process per GPU): every rank draws its own batch, the flat gradient of
each optimizer step is averaged with one all-reduce, so (1) all ranks hold identical variables afterwards and (2) the gradient
that was applied is the mean of the ranks' local gradients (computed by solo trainers on single-rank groups).
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp8x.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

import mpgan_b200  # noqa: F401
from mpgan_b200 import parallel as par, training8x as t8

rank, local, world = par.init_from_env()
dev = torch.device("cuda", local)
solo = [dist.new_group([r]) for r in range(world)][rank]           # a group of this rank alone: no averaging
mk = lambda group: t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2, learning_rate=1e-3, seed=7, lambda_t=1.0, device=local, group=group)
dp, alone = mk(None), mk(solo)
g = torch.Generator().manual_seed(100 + rank)                       # per-rank data
xs, ys = torch.rand((2, 96), generator=g).to(dev), torch.rand((2, 1024), generator=g).to(dev)
xt, yt = torch.rand((6, 96), generator=g).to(dev), torch.rand((6, 1024), generator=g).to(dev)
lf = torch.tensor([[0.3], [0.7]])
ok = True
for name, step in (("critic", lambda t: t.disc_step(xs, ys, 2.5, 2, lf)), ("temporal critic", lambda t: t.t_disc_step(xt, yt, 2.5, 2, lf)),
                   ("generator", lambda t: t.gen_step(xs, ys, 2.5, 2, xt, yt))):
    step(dp)
    step(alone)
    ps_dp = {"critic": dp.disc.ps, "temporal critic": dp.tdisc.ps, "generator": dp.gen.ps}[name]
    ps_al = {"critic": alone.disc.ps, "temporal critic": alone.tdisc.ps, "generator": alone.gen.ps}[name]
    local_g = ps_al.g.clone()
    gathered = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    mean_g = torch.stack(gathered).mean(0)
    err = float((ps_dp.g - mean_g).abs().max() / (mean_g.abs().max() + 1e-30))
    vs = [torch.empty_like(ps_dp.v) for _ in range(world)]
    dist.all_gather(vs, ps_dp.v)
    same = all(torch.equal(vs[0], v) for v in vs)
    differs = float((gathered[0] - gathered[-1]).abs().max()) > 0
    if rank == 0:
        print("%s: applied gradient vs mean of the ranks' local gradients %.2e, ranks differ locally: %s, variables identical on all ranks: %s"
              % (name, err, differs, same))
    ok = ok and err < 1e-5 and same and differs
    # keep the solo trainer on the same weights for the next step
    alone.disc.ps.v.copy_(dp.disc.ps.v); alone.tdisc.ps.v.copy_(dp.tdisc.ps.v); alone.gen.ps.v.copy_(dp.gen.ps.v)
if rank == 0:
    print("DP8X_OK" if ok else "DP8X_FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
