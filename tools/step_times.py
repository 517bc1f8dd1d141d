"""Per-launch device time of one slice batch of each 4x pass (events + sync around every step).
python tools/step_times.py [L] [precision] [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import pipeline as P, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
w1, w2 = P.make_weights_4x(L, 1)
mp = P.MultiPass4x(L, w1, w2, precision=precision, batch=batch)
x = mp.upload(synth.synthetic_volume(L, seed=1))
mp.pass1_only(x)
torch.cuda.synchronize()
st = torch.cuda.current_stream().cuda_stream
out = {}
for name, pn, inbuf in (("pass1", mp.p1.net, mp.in1), ("pass2", mp.p2.net, mp.in2)):
    for name_b, b in pn.placeholders.items():
        b.ptr = inbuf.data_ptr()
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
    tot = sum(r[1] for r in rows)
    print("== %s: %.3f ms per batch of %d slices (sum of launches), flops %.3e -> %.1f TFLOP/s" % (
        name, tot, batch, pn.flops, pn.flops / tot / 1e9))
    for i, (label, ms) in enumerate(rows):
        fl = pn.step_flops.get(i, 0.0)
        print("  %7.3f ms %5.1f%%  %7.1f TF/s  %s" % (ms, 100 * ms / tot, fl / ms / 1e9 if fl else 0.0, label))
    out[name] = rows
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/step_times_%s.json" % precision, "w"), indent=1)
