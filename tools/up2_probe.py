"""Probe of the nearest-x2-upsampling epilogue (8x net 1, 3x3 128/128->128 at 128x128 -> 256x256) vs slice batch and knock-outs.
python tools/up2_probe.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi


def run(n, hw, cins, ks, cout, ups, pn, iters=5):
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((k, k, c, cout)) * np.sqrt(2.0 / (k * k * c))).astype(np.float32) for k, c in zip(ks, cins)]
    plan = capi.ConvPlan(capi.default_handle(0), n, hw, hw, ws, cins, cout, cout, act="relu", pixel_norm=pn, upsample=ups,
                         in_dtype=capi.F16, out_dtype=capi.F16)
    xs = [torch.randn(n, hw, hw, c, device="cuda").to(torch.float16) for c in cins]
    y = torch.empty(n, hw * ups, hw * ups, cout, dtype=torch.float16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        plan.run(xs[0], xs[1] if len(xs) > 1 else None, y, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.run(xs[0], xs[1] if len(xs) > 1 else None, y, st)
    e1.record()
    torch.cuda.synchronize()
    plan.close()
    return e0.elapsed_time(e1) / iters


for dbg in ("0", "1"):
    os.environ["MPG_IGEMM_DBG"] = dbg
    for n in (8, 16, 32):
        for (ups, pn) in ((2, True), (2, False), (1, True)):
            ms = run(n, 128, [128, 128], [3, 1], 128, ups, pn)
            print("dbg=%s n=%2d ups=%d pn=%d  %.3f ms  (%.3f per 8 slices)" % (dbg, n, ups, pn, ms, ms * 8 / n), flush=True)
