"""Launch time of one thin conv vs problem size (fixed cost vs per-pixel cost). python tools/size_probe.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import mpgan_b200  # noqa
from mpgan_b200 import capi

def t(n, h, w, cins, ks, cout, iters=10):
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((k, k, c, cout)) * 0.1).astype(np.float32) for k, c in zip(ks, cins)]
    cs = [-(-c // 8) * 8 for c in cins]
    oc = -(-cout // 8) * 8
    plan = capi.ConvPlan(capi.default_handle(0), n, h, w, ws, cs, cout, oc, act="relu", in_dtype=capi.F16, out_dtype=capi.F16)
    xs = [torch.randn(n, h, w, c, device="cuda").to(torch.float16) for c in cs]
    y = torch.empty(n, h, w, oc, dtype=torch.float16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.run(xs[0], xs[1] if len(xs) > 1 else None, y, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.run(xs[0], xs[1] if len(xs) > 1 else None, y, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, plan.kind

for name, cins, ks, cout in (("cA0 4->8", [4], [5], 8), ("cA3 8->2", [8], [5], 2), ("cA1 32->128", [32], [5], 128)):
    row = []
    for n, hw in ((1, 64), (1, 256), (1, 512), (2, 512), (4, 512), (8, 512), (16, 512)):
        ms, kind = t(n, hw, hw, cins, ks, cout)
        row.append("%dx%d^2=%.3f" % (n, hw, ms))
    print(name, "kind", kind, "  ".join(row), flush=True)
