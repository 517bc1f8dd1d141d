"""Device time of the slice assembler for the two passes of the 4x recipe (CUDA events, 100 launches each).
python tools/asm_probe.py [L]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
u, B = 4, 8
S = L * u
h = capi.default_handle(0)
vol = torch.from_numpy(synth.synthetic_volume(L, seed=1)).cuda()
dens = torch.rand((S, S, S), device="cuda")
asm1 = capi.make_assemble_desc((L, L, L), 4, (0, 1, 2), (u, 1, 1), (0, 1, 2, 3), (1.0, 1.0, 1.0, 1.0), out_dtype=capi.F32, out_cstride=4)
asm2 = capi.make_assemble_desc((L, L, L), 4, (2, 0, 1), (u, u, u), (2, 3, 1), (4.0, 4.0, 4.0), out_dtype=capi.F32, out_cstride=4)
in1 = torch.empty((B, L, L, 4), device="cuda")
in2 = torch.empty((B, S, S, 4), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name, desc, d, out in (("pass1", asm1, None, in1), ("pass2", asm2, dens, in2)):
    for _ in range(5):
        capi.slice_assemble(h, desc, vol, d, 8, B, out, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(100):
        capi.slice_assemble(h, desc, vol, d, (8 * k) % (S - B), B, out, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 10.0
    byt = out.numel() * 4 + (B * S * S * 4 if d is not None else 0)
    print("%s assemble: %.1f us per batch of %d slices, %.0f GB/s of output+density bytes" % (name, us, B, byt / us / 1e3))
