"""Which stage bounds each conv of the 4x generator?  Times every layer shape of gen_resnet (8x512x512) under
pipeline-knob / knock-out settings of the igemm kernel (env is read at plan creation).
python tools/thin_probe.py [config ...]   -> gpurun_out/thin_probe.json"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi

SHAPES = [
    ("cA0 4->8", dict(cins=[4], ks=[5], cout=8)),
    ("cB0+s0 8/4->32", dict(cins=[8, 4], ks=[5, 1], cout=32)),
    ("cA1 32->128", dict(cins=[32], ks=[5], cout=128)),
    ("cB1+s1 128/32->128", dict(cins=[128, 32], ks=[5, 1], cout=128)),
    ("cA2 128->32", dict(cins=[128], ks=[5], cout=32)),
    ("cB2+s2 32/128->8", dict(cins=[32, 128], ks=[5, 1], cout=8)),
    ("cA3 8->2", dict(cins=[8], ks=[5], cout=2)),
    ("cB3+s3 2/8->1 f32", dict(cins=[2, 8], ks=[5, 1], cout=1, f32out=True)),
]
SHAPES_8X = [
    ("n2 48->96", dict(cins=[48], ks=[5], cout=96, pn=True)),
    ("n2 96/48->96", dict(cins=[96, 48], ks=[5, 1], cout=96, pn=True)),
    ("n2 96->48", dict(cins=[96], ks=[5], cout=48, pn=True)),
    ("n2 48/96->48", dict(cins=[48, 96], ks=[5, 1], cout=48, pn=True)),
    ("n2 48->48", dict(cins=[48], ks=[5], cout=48, pn=True)),
    ("n2 24/12->48", dict(cins=[24, 12], ks=[5, 1], cout=48, pn=True)),
    ("n1 64->64k3", dict(cins=[64], ks=[3], cout=64, pn=True)),
    ("n1 32->32k3", dict(cins=[32], ks=[3], cout=32, pn=True)),
    ("n2 48->24", dict(cins=[48], ks=[5], cout=24, pn=True)),
    ("n2 24->24", dict(cins=[24], ks=[5], cout=24, pn=True)),
    ("n2 5->16", dict(cins=[5], ks=[5], cout=16, pn=True)),
]
if os.environ.get("PROBE_8X"):
    SHAPES = SHAPES_8X
CONFIGS = {
    "default": {},
    "nf256": {"MPG_NFOLD_THREADS": "256"},
    # knock-outs (results are wrong, timing only): bit0 no global stores, bit1 no epilogue at all, bit2 no MMAs
    "nostore": {"MPG_IGEMM_DBG": "1", "MPG_NFOLD_DBG": "1"},
    "noepi": {"MPG_IGEMM_DBG": "2", "MPG_NFOLD_DBG": "2"},
    "nomma": {"MPG_IGEMM_DBG": "4", "MPG_NFOLD_DBG": "4"},
    "skeleton": {"MPG_IGEMM_DBG": "6", "MPG_NFOLD_DBG": "6"},
    "nobres": {"MPG_IGEMM_BRES": "0"},
    "nopair": {"MPG_IGEMM_PAIR": "0"},
    "occ1": {"MPG_IGEMM_OCC": "1"},
    "occ1_nomma": {"MPG_IGEMM_OCC": "1", "MPG_IGEMM_DBG": "4"},
    "occ1_notma": {"MPG_IGEMM_OCC": "1", "MPG_IGEMM_TMASTORE": "0"},
    "ck64": {"MPG_IGEMM_CK": "64"},
    "ck32": {"MPG_IGEMM_CK": "32"},
    "nfck32": {"MPG_NFOLD_CK": "32"},
    # row-streaming kernel (conv_vfold.cu)
    "vf_off": {"MPG_CONV_VFOLD": "0"},
    "vf_nostore": {"MPG_VFOLD_DBG": "1"},
    "vf_noepi": {"MPG_VFOLD_DBG": "2"},
    "vf_nomma": {"MPG_VFOLD_DBG": "4"},
    "vf_skel": {"MPG_VFOLD_DBG": "6"},
    "vf_na2": {"MPG_VFOLD_NA": "2"},
    "vr": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0"},
    "vr_single": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_PAIR": "0"},
    "vr_pair": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_PAIR": "1"},
    "vr_pair_noepi": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_PAIR": "1", "MPG_VRING_DBG": "2"},
    "vr_nostore": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_DBG": "1"},
    "vr_noepi": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_DBG": "2"},
    "vr_nomma": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_DBG": "4"},
    "vr_skel": {"MPG_CONV_VRING": "1", "MPG_CONV_VFOLD": "0", "MPG_VRING_DBG": "6"},
    "vf_g2": {"MPG_VFOLD_GROUPS": "2"},
    "vf_g2_nomma": {"MPG_VFOLD_GROUPS": "2", "MPG_VFOLD_DBG": "4"},
    "vf_nbuf1": {"MPG_VFOLD_NBUF": "1"},
    "nfck32_skel": {"MPG_NFOLD_CK": "32", "MPG_NFOLD_DBG": "6", "MPG_IGEMM_DBG": "6"},
}
KEYS = sorted({k for c in CONFIGS.values() for k in c})


def time_shape(sh, iters=5):
    n, h, w = 8, 512, 512
    rng = np.random.default_rng(0)
    ws = [(rng.standard_normal((k, k, c, sh["cout"])) * np.sqrt(2.0 / (k * k * c))).astype(np.float32)
          for k, c in zip(sh["ks"], sh["cins"])]
    cs = [-(-c // 8) * 8 for c in sh["cins"]]
    f32 = sh.get("f32out", False)
    oc = sh["cout"] if f32 else -(-sh["cout"] // 8) * 8
    plan = capi.ConvPlan(capi.default_handle(0), n, h, w, ws, cs, sh["cout"], oc, act="relu", in_dtype=capi.F16,
                         out_dtype=capi.F32 if f32 else capi.F16, pixel_norm=bool(sh.get("pn", False)))
    xs = [torch.randn(n, h, w, c, device="cuda").to(torch.float16) for c in cs]
    y = torch.empty(n, h, w, oc, dtype=torch.float32 if f32 else torch.float16, device="cuda")
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
    ms = e0.elapsed_time(e1) / iters
    fl = plan.flops
    plan.close()
    return ms, fl


def main():
    which = sys.argv[1:] or list(CONFIGS)
    res = {}
    for cname in which:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(CONFIGS[cname])
        row = {}
        for name, sh in SHAPES:
            try:
                ms, fl = time_shape(sh)
                row[name] = ms
            except Exception as e:  # keep the table going
                row[name] = None
                print("  %s / %s: %r" % (cname, name, e))
        res[cname] = row
        print("%-12s " % cname + "  ".join("%s=%.3f" % (k.replace(" ", "_"), v) if v is not None else "%s=ERR" % k.replace(" ", "_")
                                            for k, v in row.items()), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/thin_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
