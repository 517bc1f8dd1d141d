"""Per-launch device time of one slice batch of each generator of the shipped 8x recipe (out.py nets 1+2).
python tools/step_times_8x.py [L] [precision]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import pipeline as P, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
weights = P.make_weights_out(L, 1, upRes=8, nets=(1, 2))
mp = P.MultiPassOut(L, weights, upRes=8, precision=precision)
x = torch.from_numpy(synth.synthetic_volume(L, seed=1)).cuda()
mp(x)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
for idx in mp.nets:
    p = mp.passes[idx]
    pn = p["net"].net
    for name_b, b in pn.placeholders.items():
        b.ptr = p["inbuf"].data_ptr() if name_b == "x" else mp.vol_dim.data_ptr()
    rows = []
    for rep in range(3):
        rows = []
        for label, step in pn.steps:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            step(st)
            e1.record()
            torch.cuda.synchronize()
            rows.append((label, e0.elapsed_time(e1)))
    # whole batch, launches back to back (what the pipeline does)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        for label, step in pn.steps:
            step(st)
    e1.record()
    torch.cuda.synchronize()
    print("net %d: back-to-back %.3f ms per batch" % (idx, e0.elapsed_time(e1) / 3))
    tot = sum(r[1] for r in rows)
    print("== net %d: %.3f ms per batch of %d slices (sum of launches), flops %.3e -> %.1f TFLOP/s; %d batches per frame" % (
        idx, tot, p["batch"], pn.flops, pn.flops / tot / 1e9, mp.S // p["batch"]))
    for i, (label, ms) in enumerate(rows):
        fl = pn.step_flops.get(i, 0.0)
        print("  %7.3f ms %5.1f%%  %7.1f TF/s  %s" % (ms, 100 * ms / tot, fl / ms / 1e9 if fl else 0.0, label))
