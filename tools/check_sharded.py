"""torchrun --nproc-per-node G tools/check_sharded.py [L] [4x|8x|8x3|8x_ta1] : the slice-sharded pipelines (axis changes
as peer stores over NVLink, or NCCL all-to-all) must reproduce the single-GPU volume bit for bit."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

import mpgan_b200  # noqa: F401
from mpgan_b200 import parallel as par, pipeline as P, synth

rank, local, world = par.init_from_env()
torch.cuda.set_device(local)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = sys.argv[2] if len(sys.argv) > 2 else "4x"
x = synth.synthetic_volume(L, seed=5)
if which == "4x":
    w1, w2 = P.make_weights_4x(L, 5, randomize_bn=True)
    single = P.MultiPass4x(L, w1, w2, precision="fp16", device=local)(x).clone()
    mp = P.MultiPass4x(L, w1, w2, precision="fp16", device=local, rank=rank, world=world)
else:  # the shipped 8x recipe (GAN/example_run_output.py:18-48) with small feature counts:
    # "8x" = generators 1+2, "8x3" = all three generators (:39-47), "8x_ta1" = generators 1+2 with transposeAxis 1
    specs = {1: P.NetSpec(True, True, 64, 64, 3, True), 2: P.NetSpec(True, False, 48, 48, 5),
             3: P.NetSpec(False, False, 48, 24, 5)}
    nets = (1, 2, 3) if which == "8x3" else (1, 2)
    ta = 1 if which == "8x_ta1" else 0
    w = P.make_weights_out(L, 5, upRes=8, specs=specs, nets=nets)
    single = P.MultiPassOut(L, w, upRes=8, specs=specs, precision="fp16", device=local, transposeAxis=ta)(x).clone()
    mp = P.MultiPassOut(L, w, upRes=8, specs=specs, precision="fp16", device=local, rank=rank, world=world, transposeAxis=ta)
part = mp(x)
part = mp(x)  # a second frame through the same buffers (slab reuse across frames)
torch.cuda.synchronize()
ref = single[mp.s0:mp.s1]
same = bool(torch.equal(part, ref))
maxd = float((part - ref).abs().max())
flag = torch.tensor([1 if same else 0], device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
print("rank %d/%d slab [%d,%d): bit-exact=%s max|d|=%.3e  nonzero=%.3f" % (rank, world, mp.s0, mp.s1, same, maxd,
                                                                        float((part != 0).float().mean())))
if rank == 0:
    print("SHARDED_OK" if int(flag.item()) == 1 else "SHARDED_MISMATCH")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
