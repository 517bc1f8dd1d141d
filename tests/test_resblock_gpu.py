"""GPU parity of the fused thin residual block (mpg_resblock_plan_*, csrc/resblock_thin.cu) through the C ABI against a
plain torch fp64 reference of the same three convolutions (GAN/multipassGAN-4x.py:505-526) on the same 16-bit-rounded
operands: ru1 (4->8->32, fp32 rows, optional nearest x4 input) and ru4 (8->2->1, 16-bit input, fp32 output) of
gen_resnet plus ragged sizes, lrelu / no activation, BN scales and the padded-channel contract."""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi
from convref import ref_conv

pytestmark = pytest.mark.gpu

_TDT = {"bf16": torch.bfloat16, "f16": torch.float16}
_CODE = {"bf16": capi.BF16, "f16": capi.F16, "f32": capi.F32}

CASES = {
    "ru1_pass2": dict(n=2, h=64, w=96, cin=4, cmid=8, cout=32, in_f32=True),
    "ru1_pass1_up4": dict(n=2, h=64, w=64, cin=4, cmid=8, cout=32, in_f32=True, up=4),
    "ru1_ragged_37x45": dict(n=3, h=37, w=45, cin=4, cmid=8, cout=32, in_f32=True, act="lrelu"),
    "ru1_cin3_cmid5": dict(n=1, h=40, w=33, cin=3, cmid=5, cout=32, in_f32=True, act=None),
    "ru1_tiny_3x5": dict(n=1, h=3, w=5, cin=4, cmid=8, cout=32, in_f32=True),
    "ru1_16bit_in": dict(n=1, h=48, w=40, cin=8, cmid=8, cout=32, in_f32=False),
    "ru4": dict(n=2, h=64, w=96, cin=8, cmid=2, cout=1, in_f32=False, out_f32=True),
    "ru4_ragged": dict(n=3, h=37, w=45, cin=8, cmid=2, cout=1, in_f32=False, out_f32=True, act="lrelu"),
    "thin_8_8_8_16bit_out": dict(n=1, h=33, w=70, cin=6, cmid=8, cout=7, in_f32=False, out_f32=False),
    "thin_f32_out_cs4": dict(n=1, h=32, w=32, cin=8, cmid=4, cout=3, in_f32=False, out_f32=True, out_cstride=4),
}


@pytest.mark.parametrize("half", ["f16", "bf16"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_resblock_case(name, half):
    c = dict(CASES[name])
    n, h, w, cin, cmid, cout = c["n"], c["h"], c["w"], c["cin"], c["cmid"], c["cout"]
    up, act = c.get("up", 1), c.get("act", "relu")
    in_f32, out_f32 = c["in_f32"], c.get("out_f32", False)
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7)
    tdt = _TDT[half]
    in_cs = 4 if in_f32 else 8
    out_cs = c.get("out_cstride", cout if out_f32 else -(-cout // 8) * 8)
    x = torch.randn(n, h // up if up > 1 else h, w // up if up > 1 else w, in_cs, generator=g)
    x[..., cin:] = 0.0 if in_f32 else 3.0  # fp32 rows: unused channels are zero; 16-bit: garbage must be ignored via zero weights
    xd = x.to(dev) if in_f32 else x.to(tdt).to(dev)
    xv = xd.float().to(tdt).float()[..., :cin]  # the values the MMAs see
    std = lambda k, ci: np.sqrt(2.0) / np.sqrt(k * k * ci)
    wa = (torch.randn(5, 5, cin, cmid, generator=g) * std(5, cin)).numpy()
    wb = (torch.randn(5, 5, cmid, cout, generator=g) * std(5, cmid)).numpy()
    ws = (torch.randn(1, 1, cin, cout, generator=g) * std(1, cin)).numpy()
    sca, scb, scs = [(0.5 + torch.rand(k, generator=g)).numpy() for k in (cmid, cout, cout)]
    sha = (torch.randn(cmid, generator=g) * 0.1).numpy()
    shbs = (torch.randn(cout, generator=g) * 0.1).numpy()
    plan = capi.ResblockPlan(capi.default_handle(0), n, h, w, wa, wb, ws, capi.F32 if in_f32 else _CODE[half], in_cs,
                             _CODE[half], capi.F32 if out_f32 else _CODE[half], out_cs, act=act, scale_a=sca, scale_b=scb,
                             scale_s=scs, shift_a=sha, shift_bs=shbs, in_upsample=up)
    y = torch.full((n, h, w, out_cs), float("nan"), dtype=torch.float32 if out_f32 else tdt, device=dev)
    plan.run(xd, y, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    plan.close()

    mid = ref_conv([xv], [wa], [sca], sha, act, False, 1, in_upsample=up, round_w=tdt)
    mid = mid.float().to(tdt).double()  # the intermediate is stored as 16-bit in shared memory
    xu = xv.repeat_interleave(up, 1).repeat_interleave(up, 2) if up > 1 else xv  # the shortcut reads the upsampled view
    ref = ref_conv([mid, xu.double()], [wb, ws], [scb, scs], shbs, act, False, 1, round_w=tdt)
    got = y.double()
    assert torch.isfinite(got).all()
    if out_cs > cout:
        assert (got[..., cout:] == 0).all(), "padded output channels must be written as zeros"
    d = (got[..., :cout] - ref).abs()
    rel = float(torch.linalg.norm(d) / (torch.linalg.norm(ref) + 1e-30))
    tol = 2e-5 if out_f32 else (6e-3 if half == "bf16" else 8e-4)
    # the 16-bit rounding of the intermediate can flip by one ulp against the fp64 reference: allow that through rel-L2
    if out_f32:
        tol = 4e-3 if half == "bf16" else 5e-4
    print(name, half, "rel_l2 %.3e max %.3e ref_max %.3f" % (rel, float(d.max()), float(ref.abs().max())))
    assert rel < tol, (name, half, rel)


def test_resblock_unsupported_is_loud():
    hd = capi.default_handle(0)
    z = lambda *s: np.zeros(s, np.float32)
    with pytest.raises(capi.MpgError):  # cmid > 8
        capi.ResblockPlan(hd, 1, 32, 32, z(5, 5, 4, 16), z(5, 5, 16, 32), z(1, 1, 4, 32), capi.F32, 4, capi.F16, capi.F16, 32)
    with pytest.raises(capi.MpgError):  # 3x3
        capi.ResblockPlan(hd, 1, 32, 32, z(3, 3, 4, 8), z(3, 3, 8, 32), z(1, 1, 4, 32), capi.F32, 4, capi.F16, capi.F16, 32)
    with pytest.raises(capi.MpgError):  # 16 output channels: neither the 32-channel nor the <= 8-channel layout
        capi.ResblockPlan(hd, 1, 32, 32, z(5, 5, 4, 8), z(5, 5, 8, 16), z(1, 1, 4, 16), capi.F32, 4, capi.F16, capi.F16, 16)
