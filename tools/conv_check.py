"""GPU dev check: sweep conv cases through the C ABI, print error stats, then time the flagship shapes.
Run on the GPU box:  python tools/conv_check.py [--perf]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

from convref import run_case
import mpgan_b200
from mpgan_b200 import capi

CASES = [
    # name, kwargs
    ("k1_64to128", dict(n=1, h=32, w=32, cins=[64], ks=[1], cout=128)),
    ("k1_64to16", dict(n=1, h=16, w=16, cins=[64], ks=[1], cout=16)),
    ("k3_64to64", dict(n=1, h=32, w=32, cins=[64], ks=[3], cout=64)),
    ("k5_128to128", dict(n=1, h=64, w=64, cins=[128], ks=[5], cout=128, act="relu")),
    ("k5_128to32", dict(n=2, h=48, w=48, cins=[128], ks=[5], cout=32, act="relu")),
    ("k5_32to128_ck32", dict(n=1, h=64, w=64, cins=[32], ks=[5], cout=128, act="relu")),
    ("k5_16to32_ck16", dict(n=1, h=64, w=64, cins=[16], ks=[5], cout=32, act="relu")),
    ("k5_8to32_ck16", dict(n=1, h=32, w=32, cins=[8], ks=[5], cout=32, cstrides=[16], act="relu")),
    ("2seg_128k5_32k1", dict(n=1, h=64, w=64, cins=[128, 32], ks=[5, 1], cout=128, act="relu")),
    ("2seg_32k5_128k1_to8", dict(n=1, h=64, w=64, cins=[32, 128], ks=[5, 1], cout=8, act="relu", out_cstride=16)),
    ("k3_pn_up2", dict(n=1, h=32, w=32, cins=[128], ks=[3], cout=128, act="relu", pixel_norm=True, upsample=2)),
    ("ragged_37x45", dict(n=3, h=37, w=45, cins=[64], ks=[5], cout=48, act="lrelu")),
    ("k5_f32out", dict(n=1, h=32, w=32, cins=[128], ks=[5], cout=24, out_dtype="f32")),
    ("k3_96to96", dict(n=1, h=40, w=40, cins=[96], ks=[3], cout=96, act="relu", pixel_norm=True)),
    ("big_512", dict(n=1, h=512, w=512, cins=[128], ks=[5], cout=128, act="relu")),
    # CUDA-core direct path
    ("direct_f32_k5_4to8_up4", dict(n=2, h=64, w=64, cins=[4], ks=[5], cout=8, in_dtype="f32", out_dtype="bf16", in_upsample=4, act="relu")),
    ("direct_bf16_k5_8to2", dict(n=1, h=40, w=40, cins=[8], ks=[5], cout=2, cstrides=[16], out_cstride=8, act="relu")),
    ("direct_2seg_to1_f32", dict(n=1, h=40, w=40, cins=[2, 8], ks=[5, 1], cout=1, cstrides=[8, 16], out_dtype="f32", act="relu")),
    ("direct_f32_k4s2", dict(n=2, h=64, w=64, cins=[2], ks=[4], cout=32, in_dtype="f32", out_dtype="f32", stride=2, act="lrelu")),
    ("direct_f32_k4s1", dict(n=2, h=8, w=8, cins=[16], ks=[4], cout=24, in_dtype="f32", out_dtype="f32", stride=1, act="lrelu")),
    ("direct_f32_128to128_pn", dict(n=1, h=24, w=24, cins=[128, 6], ks=[3, 1], cout=128, in_dtype="f32", out_dtype="f32", act="relu", pixel_norm=True, upsample=2)),
    ("forced_direct_bf16", dict(n=1, h=32, w=32, cins=[64], ks=[3], cout=64, force_kind=2, act="relu")),
]


def perf(n, h, w, cins, ks, cout, iters=5, **kw):
    dev = torch.device("cuda:0")
    import numpy as np
    ws = [np.random.randn(k, k, c, cout).astype("float32") * 0.02 for k, c in zip(ks, cins)]
    plan = capi.ConvPlan(capi.default_handle(0), n, h, w, ws, cins, cout, cout, act="relu", **kw)
    xs = [torch.randn(n, h, w, c, device=dev).to(torch.bfloat16) for c in cins]
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
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
    ms = e0.elapsed_time(e1) / iters
    return dict(ms=ms, tflops=plan.flops / ms / 1e9, kind=plan.kind)


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    out = []
    bad = 0
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, kw in CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        try:
            r = run_case(**kw)
        except Exception as e:  # keep going: we want the whole table from one GPU call
            r = dict(error=repr(e))
        r["name"] = name
        r["sec"] = round(time.time() - t0, 2)
        tol = 6e-3 if kw.get("out_dtype", "bf16") == "bf16" else 2e-5
        ok = ("error" not in r) and r["finite"] and r["pad_ok"] and r["rel_l2"] < tol
        r["ok"] = ok
        bad += (not ok)
        print(json.dumps(r), flush=True)
        out.append(r)
    if "--perf" in sys.argv:
        for name, kw in [
            ("ru2_B 8x512^2 128->128 k5", dict(n=8, h=512, w=512, cins=[128], ks=[5], cout=128)),
            ("ru2_B+s 8x512^2", dict(n=8, h=512, w=512, cins=[128, 32], ks=[5, 1], cout=128)),
            ("ru2_A 8x512^2 32->128 k5", dict(n=8, h=512, w=512, cins=[32], ks=[5], cout=128)),
            ("ru3_A 8x512^2 128->32 k5", dict(n=8, h=512, w=512, cins=[128], ks=[5], cout=32)),
            ("net1 8x128^2 128->128 k3", dict(n=8, h=128, w=128, cins=[128], ks=[3], cout=128)),
            ("ru1_B 8x512^2 16->32 k5", dict(n=8, h=512, w=512, cins=[16], ks=[5], cout=32)),
        ]:
            try:
                r = perf(**kw)
            except Exception as e:
                r = dict(error=repr(e))
            r["name"] = name
            print(json.dumps(r), flush=True)
            out.append(r)
    with open("gpurun_out/conv_check.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print("FAILED cases:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
