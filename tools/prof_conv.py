"""Run conv shapes a few times each (target of the ncu --set full capture).
python tools/prof_conv.py [ru2b|ru2a|ru3a|ru1a|n2_48to48|n2_96to48|n2_48_96to48|n1_64to64k3][,more] [iters]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi

SHAPES = {
    "ru2b": dict(cins=[128, 32], ks=[5, 1], cout=128),
    "ru2a": dict(cins=[32], ks=[5], cout=128),
    "ru3a": dict(cins=[128], ks=[5], cout=32),
    "ru1a": dict(cins=[4], ks=[5], cout=8),
    # layers of the 8x generators (pixel_norm): the TMEM-ring row-streaming kernel, single CTA / CTA pairs
    "n2_48to48": dict(cins=[48], ks=[5], cout=48, pn=True),
    "n2_96to48": dict(cins=[96], ks=[5], cout=48, pn=True),
    "n2_48_96to48": dict(cins=[48, 96], ks=[5, 1], cout=48, pn=True),
    "n1_64to64k3": dict(cins=[64], ks=[3], cout=64, pn=True),
}
names = (sys.argv[1] if len(sys.argv) > 1 else "ru2b").split(",")
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n, h, w = 8, 512, 512
for name in names:
    sh = SHAPES[name]
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((k, k, c, sh["cout"])) * np.sqrt(2.0 / (k * k * c))).astype(np.float32)
          for k, c in zip(sh["ks"], sh["cins"])]
    cs = [-(-c // 8) * 8 for c in sh["cins"]]
    oc = -(-sh["cout"] // 8) * 8
    plan = capi.ConvPlan(capi.default_handle(0), n, h, w, ws, cs, sh["cout"], oc, act="relu", in_dtype=capi.F16,
                         out_dtype=capi.F16, pixel_norm=bool(sh.get("pn", False)))
    xs = [torch.randn(n, h, w, c, device="cuda").to(torch.float16) for c in cs]
    y = torch.empty(n, h, w, oc, dtype=torch.float16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(iters):
        if i == iters - 1:
            e0.record()
        plan.run(xs[0], xs[1] if len(xs) > 1 else None, y, st)
    e1.record()
    torch.cuda.synchronize()
    print(name, "kind", plan.kind, "last launch ms", e0.elapsed_time(e1), "TFLOP/s", plan.flops / e0.elapsed_time(e1) / 1e9)
    plan.close()
