"""Shared helpers for the GPU conv parity tests: build a ConvPlan case, run it through the C ABI and
compare with a plain fp32 torch reference of the same op (same bf16-rounded operands)."""
import numpy as np
import torch
import torch.nn.functional as F

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi


_TDT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}
_CODE = {"bf16": capi.BF16, "f16": capi.F16, "f32": capi.F32}


def tf_same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def ref_conv(xs, ws, scales, shift, act, pixel_norm, upsample, in_upsample=1, stride=1, round_w=None):
    """xs: list of NHWC fp32 tensors (already holding the values the kernel sees); ws: HWIO fp32."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    acc = None
    for x, w, sc in zip(xs, ws, scales):
        w = torch.as_tensor(w, dtype=torch.float32, device=x.device)
        if sc is not None:
            w = w * torch.as_tensor(sc, dtype=torch.float32, device=x.device)
        if round_w is not None:
            w = w.to(round_w).to(torch.float32)
        xi = x.permute(0, 3, 1, 2)
        if in_upsample > 1:
            xi = xi.repeat_interleave(in_upsample, 2).repeat_interleave(in_upsample, 3)
        k = w.shape[0]
        pt, pb = tf_same_pad(xi.shape[2], k, stride)
        pl, pr = tf_same_pad(xi.shape[3], k, stride)
        xi = F.pad(xi.double(), (pl, pr, pt, pb))
        y = F.conv2d(xi, w.permute(3, 2, 0, 1).double(), stride=stride)
        acc = y if acc is None else acc + y
    if shift is not None:
        acc = acc + torch.as_tensor(shift, dtype=torch.float64, device=acc.device).view(1, -1, 1, 1)
    if act == "relu":
        acc = torch.relu(acc)
    elif act == "lrelu":
        acc = 0.6 * acc + 0.4 * acc.abs()
    elif act == "tanh":
        acc = torch.tanh(acc)
    if pixel_norm:
        acc = acc * torch.rsqrt((acc * acc).mean(dim=1, keepdim=True) + 1e-8)
    if upsample > 1:
        acc = acc.repeat_interleave(upsample, 2).repeat_interleave(upsample, 3)
    return acc.permute(0, 2, 3, 1).contiguous()


def run_case(n, h, w, cins, ks, cout, act=None, pixel_norm=False, upsample=1, in_upsample=1, stride=1,
             in_dtype="bf16", out_dtype="bf16", force_kind=0, seed=0, cstrides=None, out_cstride=None,
             with_scale=True):
    """Returns dict(kind, max_abs, rel_l2, ref_max)."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    nseg = len(cins)
    if cstrides is None:
        cstrides = [(-(-c // 8) * 8) if in_dtype != "f32" else c for c in cins]
    if out_cstride is None:
        out_cstride = (-(-cout // 8) * 8) if out_dtype != "f32" else cout
    sh, sw = h // in_upsample, w // in_upsample
    xs_dev, xs_val, ws, scs = [], [], [], []
    for s in range(nseg):
        x = torch.randn(n, sh, sw, cstrides[s], generator=g)
        if in_dtype != "f32":
            xd = x.to(_TDT[in_dtype]).to(dev)
            xv = xd.float()[..., :cins[s]]
        else:
            xd = x.to(dev)
            xv = xd[..., :cins[s]]
        xs_dev.append(xd.contiguous())
        xs_val.append(xv)
        wt = torch.randn(ks[s], ks[s], cins[s], cout, generator=g) * (np.sqrt(2.0) / np.sqrt(ks[s] * ks[s] * cins[s]))
        ws.append(wt.numpy())
        scs.append((0.5 + torch.rand(cout, generator=g)).numpy() if with_scale else None)
    shift = torch.randn(cout, generator=g).numpy() * 0.1
    plan = capi.ConvPlan(capi.default_handle(0), n, h, w, ws, cstrides, cout, out_cstride, act=act,
                         scales=scs if with_scale else None, shift=shift, pixel_norm=pixel_norm,
                         upsample=upsample, in_upsample=in_upsample, stride=stride,
                         in_dtype=_CODE[in_dtype], out_dtype=_CODE[out_dtype], force_kind=force_kind)
    y = torch.full((n, plan.out_h, plan.out_w, out_cstride), float("nan"),
                   dtype=_TDT[out_dtype], device=dev)
    plan.run(xs_dev[0], xs_dev[1] if nseg > 1 else None, y, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = ref_conv(xs_val, ws, scs, shift, act, pixel_norm, upsample, in_upsample, stride,
                   round_w=_TDT[in_dtype] if plan.kind in (capi.KIND_TCGEN05, capi.KIND_NFOLD, capi.KIND_VFOLD, capi.KIND_VRING) else None)
    got = y.double()
    pad_ok = True
    if out_cstride > cout:
        pad_ok = bool((got[..., cout:] == 0).all())
    got = got[..., :cout]
    diff = (got - ref).abs()
    res = dict(kind=plan.kind, max_abs=float(diff.max()), ref_max=float(ref.abs().max()),
               rel_l2=float(torch.linalg.norm(diff) / (torch.linalg.norm(ref) + 1e-30)),
               finite=bool(torch.isfinite(got).all()), pad_ok=pad_ok, flops=plan.flops)
    plan.close()
    return res
