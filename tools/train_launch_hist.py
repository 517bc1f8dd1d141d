"""Histogram of the launches of one training loop body (eager, precision fp16): which kernels make up the launch storm."""
import collections
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi, training

hist = collections.Counter()
tms = collections.Counter()
TIMED = os.environ.get("TIMED", "1") == "1"


def timed(key, fn):
    if not TIMED:
        return fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    tms[key] += e0.elapsed_time(e1)
    return r

orig_call, orig_pack = capi.train_call, capi.pack_channels
orig_run, orig_upd = capi.ConvPlan.run, capi.ConvPlan.update


def tc(name, *a):
    hist[name] += 1
    key = name
    if name in ("conv_fwd", "conv_dgrad", "conv_wgrad"):
        key = "%s k%d s%d %d->%d @%d" % (name, a[-4] if name != "conv_dgrad" else a[-4], a[-3] if name != "conv_dgrad" else a[-3], a[-6] if name != "conv_dgrad" else a[-6], a[-5] if name != "conv_dgrad" else a[-5], a[-8] if name != "conv_dgrad" else a[-8])
    return timed(key, lambda: orig_call(name, *a))


def pk(*a, **k):
    hist["pack_channels"] += 1
    srcs = a[1]
    key = "pack %dsrc f%d %s -> %s cs%d @%dx%dx%d" % (len(srcs), srcs[0][5], "+".join(str(t[4]) for t in srcs),
                                                    {0: "bf16", 1: "f16", 2: "f32"}.get(a[3], a[3]), a[4], a[5], a[6], a[7])
    return timed(key, lambda: orig_pack(*a, **k))


def run(self, *a, **k):
    hist["conv_plan_run"] += 1
    return timed("conv_plan_run", lambda: orig_run(self, *a, **k))


def upd(self, *a, **k):
    hist["conv_plan_update"] += 1
    return timed("conv_plan_update", lambda: orig_upd(self, *a, **k))


capi.train_call, capi.pack_channels = tc, pk
capi.ConvPlan.run, capi.ConvPlan.update = run, upd
tr = training.Trainer4x(batch=16, precision="fp16", graphs=False)
rng = np.random.default_rng(0)
x = rng.random((16, 16 * 16 * 4), dtype=np.float32)
y = rng.random((16, 64 * 64), dtype=np.float32)
tr.iteration([(x, y)], [(x, y)])
hist.clear()
tms.clear()
tr.iteration([(x, y)], [(x, y)])
torch.cuda.synchronize()
tot = sum(hist.values())
print("launch-issuing calls per loop body:", tot)
for k, v in hist.most_common():
    print("%6d %5.1f%%  %s" % (v, 100.0 * v / tot, k))
tt = sum(tms.values())
print("device time per loop body (events around every call, synchronised): %.3f ms" % tt)
for k, v in tms.most_common(40):
    print("%8.3f ms %5.1f%%  %s" % (v, 100.0 * v / tt, k))
