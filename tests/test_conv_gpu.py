"""GPU parity of the fused convolution kernels (through the C ABI) vs a plain fp32/fp64 torch reference."""
import pytest

import mpgan_b200  # noqa: F401
from convref import run_case

pytestmark = pytest.mark.gpu

CASES = {
    "k1_64to128": dict(n=1, h=32, w=32, cins=[64], ks=[1], cout=128),
    "k3_64to64": dict(n=1, h=32, w=32, cins=[64], ks=[3], cout=64),
    "k5_128to128": dict(n=1, h=64, w=64, cins=[128], ks=[5], cout=128, act="relu"),
    "k5_128to32": dict(n=2, h=48, w=48, cins=[128], ks=[5], cout=32, act="relu"),
    "k5_32to128_ck32": dict(n=1, h=64, w=64, cins=[32], ks=[5], cout=128, act="relu"),
    "k5_16to32_ck16": dict(n=1, h=64, w=64, cins=[16], ks=[5], cout=32, act="relu"),
    "k5_8to32_4s": dict(n=1, h=32, w=32, cins=[8, 4], ks=[5, 1], cout=32, cstrides=[8, 8], act="relu"),
    "2seg_128k5_32k1": dict(n=1, h=64, w=64, cins=[128, 32], ks=[5, 1], cout=128, act="relu"),
    "2seg_32k5_128k1_to8": dict(n=1, h=64, w=64, cins=[32, 128], ks=[5, 1], cout=8, act="relu"),
    "k3_pn_up2": dict(n=1, h=32, w=32, cins=[128], ks=[3], cout=128, act="relu", pixel_norm=True, upsample=2),
    # nearest x2 store: bulk tensor stores through the 5-D map (whole tiles per image, several images) / per-lane fallback
    "k3_up2_tma_n3": dict(n=3, h=32, w=48, cins=[128, 64], ks=[3, 1], cout=128, act="relu", pixel_norm=True, upsample=2),
    "k3_up2_ragged_h24": dict(n=2, h=24, w=40, cins=[128], ks=[3], cout=128, act="relu", pixel_norm=True, upsample=2),
    "k3_up2_64ch": dict(n=2, h=32, w=32, cins=[64, 64], ks=[3, 1], cout=64, act="relu", pixel_norm=True, upsample=2),
    "ragged_37x45": dict(n=3, h=37, w=45, cins=[64], ks=[5], cout=48, act="lrelu"),
    "tiny_3x5": dict(n=1, h=3, w=5, cins=[16], ks=[3], cout=16, act="tanh"),
    "nf_tiny_3x5": dict(n=1, h=3, w=5, cins=[16], ks=[3], cout=16, act="relu", force_kind=3),
    "k5_f32out": dict(n=1, h=32, w=32, cins=[128], ks=[5], cout=24, out_dtype="f32"),
    "k5_96to48": dict(n=1, h=40, w=40, cins=[96, 48], ks=[5, 1], cout=48, act="relu", pixel_norm=True),
    "k5_24to12": dict(n=1, h=40, w=40, cins=[24], ks=[5], cout=12, act="relu", pixel_norm=True),
    # tap-folded tcgen05 path (narrow Cout): every thin layer of gen_resnet + pixel-norm / k3 / ragged variants
    "nf_k5_4to8": dict(n=1, h=64, w=64, cins=[4], ks=[5], cout=8, act="relu", force_kind=3),
    "nf_k5_8to2": dict(n=2, h=40, w=40, cins=[8], ks=[5], cout=2, act="relu", force_kind=3),
    "nf_2seg_2k5_8k1_to1_f32": dict(n=1, h=40, w=72, cins=[2, 8], ks=[5, 1], cout=1, out_dtype="f32", act="relu", force_kind=3),
    "nf_k3_32to32_pn": dict(n=1, h=48, w=48, cins=[32], ks=[3], cout=32, act="lrelu", pixel_norm=True, force_kind=3),
    "nf_k3_64to32_s": dict(n=1, h=32, w=64, cins=[64, 64], ks=[3, 1], cout=32, act="relu", pixel_norm=True, force_kind=3),
    "nf_ragged_37x45": dict(n=3, h=37, w=45, cins=[48], ks=[5], cout=24, act="lrelu", force_kind=3),
    "nf_ns8_8and4_to32": dict(n=2, h=40, w=72, cins=[8, 4], ks=[5, 1], cout=32, act="relu", force_kind=3),
    "nf_ns8_k3_8to8_pn": dict(n=1, h=33, w=47, cins=[8], ks=[3], cout=8, act="lrelu", pixel_norm=True, force_kind=3),
    "nf_ns8_narrow_w20": dict(n=1, h=12, w=20, cins=[4], ks=[5], cout=8, act="relu", force_kind=3),
    "nf_k5_128to32_256": dict(n=1, h=256, w=256, cins=[128], ks=[5], cout=32, act="relu", force_kind=3),
    "direct_f32_k5_4to8_up4": dict(n=2, h=64, w=64, cins=[4], ks=[5], cout=8, in_dtype="f32", in_upsample=4, act="relu"),
    "direct_k5_8to2": dict(n=1, h=40, w=40, cins=[8], ks=[5], cout=2, act="relu"),
    "direct_2seg_to1_f32": dict(n=1, h=40, w=40, cins=[2, 8], ks=[5, 1], cout=1, out_dtype="f32", act="relu"),
    # tap-folded CTA-pair configuration (resident half weight tiles, one accumulator per tile, 3 TMEM buffers)
    "nfp_k5_128to32_odd_tiles": dict(n=1, h=12, w=28, cins=[128], ks=[5], cout=32, act="relu", force_kind=3),
    "nfp_k5_128to32_ragged": dict(n=3, h=37, w=45, cins=[128], ks=[5], cout=32, act="lrelu", force_kind=3),
    "nfp_k3_128to32_pn": dict(n=2, h=40, w=72, cins=[128], ks=[3], cout=32, act="relu", pixel_norm=True, force_kind=3),
    "nfp_k5_96and32_to24_f32": dict(n=1, h=33, w=64, cins=[96, 32], ks=[5, 1], cout=24, out_dtype="f32", act="relu", force_kind=3),
    "direct_f32_k4s2": dict(n=2, h=64, w=64, cins=[2], ks=[4], cout=32, in_dtype="f32", out_dtype="f32", stride=2, act="lrelu"),
    "direct_f32_k4s1": dict(n=2, h=8, w=8, cins=[16], ks=[4], cout=24, in_dtype="f32", out_dtype="f32", stride=1, act="lrelu"),
    "direct_f32_128_pn_up2": dict(n=1, h=24, w=24, cins=[128, 6], ks=[3, 1], cout=128, in_dtype="f32", out_dtype="f32",
                                  act="relu", pixel_norm=True, upsample=2),
    # CUDA-core kernel for the channel-less tail layers (conv_tiny.cu): ru3 of gen_resnet + ragged / k3 / lrelu variants
    "ct_k5_8to2": dict(n=2, h=40, w=72, cins=[8], ks=[5], cout=2, act="relu", force_kind=4),
    "ct_2seg_2k5_8k1_to1_f32": dict(n=2, h=40, w=136, cins=[2, 8], ks=[5, 1], cout=1, out_dtype="f32", act="relu", force_kind=4),
    "ct_ragged_37x45_k3_4to2": dict(n=3, h=37, w=45, cins=[4, 5], ks=[3, 1], cout=2, act="lrelu", with_scale=True, force_kind=4),
    "ct_k5_3to1_16bit_out": dict(n=1, h=19, w=130, cins=[3], ks=[5], cout=1, act="tanh", force_kind=4),
    "ct_k3_8to1_f32_cs4": dict(n=1, h=16, w=64, cins=[8], ks=[3], cout=1, out_dtype="f32", out_cstride=4, force_kind=4),
    # row-streaming tcgen05 path (conv_vfold.cu): vertical taps folded into N, CTA pairs on adjacent 128-pixel strips
    "vf_k5_128to32": dict(n=2, h=40, w=300, cins=[128], ks=[5], cout=32, act="relu", force_kind=5),
    "vf_k5_48to48_pn": dict(n=1, h=33, w=256, cins=[48], ks=[5], cout=48, act="relu", pixel_norm=True, force_kind=5),
    "vf_k5_48and96_to48_pn": dict(n=2, h=24, w=140, cins=[48, 96], ks=[5, 1], cout=48, act="relu", pixel_norm=True, force_kind=5),
    "vf_k5_96to48_ck32": dict(n=1, h=20, w=260, cins=[96], ks=[5], cout=48, act="lrelu", force_kind=5),
    "vf_k5_24and48_to24_pn": dict(n=1, h=30, w=130, cins=[24, 48], ks=[5, 1], cout=24, act="relu", pixel_norm=True, force_kind=5),
    "vf_k5_24to12_pn": dict(n=3, h=17, w=64, cins=[24], ks=[5], cout=12, act="relu", pixel_norm=True, force_kind=5),
    "vf_k5_32to8": dict(n=2, h=37, w=45, cins=[32], ks=[5], cout=8, act="relu", force_kind=5),
    "vf_k5_32and128_to8": dict(n=1, h=21, w=129, cins=[32, 128], ks=[5, 1], cout=8, act="relu", force_kind=5),
    "vf_k3_64to64_pn": dict(n=2, h=32, w=200, cins=[64], ks=[3], cout=64, act="relu", pixel_norm=True, force_kind=5),
    "vf_k3_64and64_to32_pn": dict(n=1, h=48, w=256, cins=[64, 64], ks=[3, 1], cout=32, act="relu", pixel_norm=True, force_kind=5),
    "vf_k3_32and32_to24_tanh": dict(n=1, h=19, w=77, cins=[32, 32], ks=[3, 1], cout=24, act="tanh", force_kind=5),
    "vf_k5_128to24_f32": dict(n=1, h=16, w=131, cins=[128], ks=[5], cout=24, out_dtype="f32", act="relu", force_kind=5),
    "vf_ragged_37x45": dict(n=3, h=37, w=45, cins=[48], ks=[5], cout=24, act="lrelu", with_scale=True, force_kind=5),
    "vf_tiny_3x5": dict(n=1, h=3, w=5, cins=[16], ks=[3], cout=16, act="relu", force_kind=5),
    "vf_k5_128to32_512": dict(n=2, h=128, w=512, cins=[128], ks=[5], cout=32, act="relu", force_kind=5),
    # row-streaming tcgen05 path with the vertical-tap sum accumulated in a TMEM ring (conv_vring.cu)
    "vr_k5_48to48_pn": dict(n=1, h=33, w=256, cins=[48], ks=[5], cout=48, act="relu", pixel_norm=True, force_kind=6),
    "vr_k5_48and96_to48_pn": dict(n=2, h=24, w=140, cins=[48, 96], ks=[5, 1], cout=48, act="relu", pixel_norm=True, force_kind=6),
    "vr_k5_24and48_to24_pn": dict(n=1, h=30, w=130, cins=[24, 48], ks=[5, 1], cout=24, act="relu", pixel_norm=True, force_kind=6),
    "vr_k5_24to12_pn": dict(n=3, h=17, w=64, cins=[24], ks=[5], cout=12, act="relu", pixel_norm=True, force_kind=6),
    "vr_k5_5to16_pn": dict(n=2, h=40, w=200, cins=[5], ks=[5], cout=16, act="relu", pixel_norm=True, force_kind=6),
    "vr_k5_32to8": dict(n=2, h=37, w=45, cins=[32], ks=[5], cout=8, act="relu", force_kind=6),
    "vr_k5_32and128_to8": dict(n=1, h=21, w=129, cins=[32, 128], ks=[5, 1], cout=8, act="relu", force_kind=6),
    "vr_k3_64to64_pn": dict(n=2, h=32, w=200, cins=[64], ks=[3], cout=64, act="relu", pixel_norm=True, force_kind=6),
    "vr_k3_64and64_to32_pn": dict(n=1, h=48, w=256, cins=[64, 64], ks=[3, 1], cout=32, act="relu", pixel_norm=True, force_kind=6),
    "vr_k3_32and32_to24_tanh": dict(n=1, h=19, w=77, cins=[32, 32], ks=[3, 1], cout=24, act="tanh", force_kind=6),
    "vr_k5_64to24_f32": dict(n=1, h=16, w=131, cins=[64], ks=[5], cout=24, out_dtype="f32", act="relu", force_kind=6),
    "vr_ragged_37x45": dict(n=3, h=37, w=45, cins=[48], ks=[5], cout=24, act="lrelu", with_scale=True, force_kind=6),
    "vr_tiny_3x5": dict(n=1, h=3, w=5, cins=[16], ks=[3], cout=16, act="relu", force_kind=6),
    "vr_k5_96to48_pair_only": dict(n=1, h=20, w=260, cins=[96], ks=[5], cout=48, act="lrelu", force_kind=6),
    "vr_k5_128to32_pair_only": dict(n=2, h=40, w=300, cins=[128], ks=[5], cout=32, act="relu", force_kind=6),
    "vr_k5_48to48_512": dict(n=2, h=128, w=512, cins=[48], ks=[5], cout=48, act="relu", pixel_norm=True, force_kind=6),
    "forced_direct_16bit": dict(n=1, h=32, w=32, cins=[64], ks=[3], cout=64, force_kind=2, act="relu"),
}


@pytest.mark.parametrize("half", ["bf16", "f16"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_conv_case(name, half):
    kw = dict(CASES[name])
    if kw.get("in_dtype", "bf16") == "f32" and half == "f16" and kw.get("out_dtype", "bf16") == "f32":
        pytest.skip("pure fp32 case does not depend on the 16-bit type")
    for key in ("in_dtype", "out_dtype"):
        if kw.get(key, "bf16") == "bf16":
            kw[key] = half
    r = run_case(**kw)
    out16 = kw["out_dtype"] != "f32"
    tol = (6e-3 if half == "bf16" else 8e-4) if out16 else 2e-5
    assert r["finite"] and r["pad_ok"], r
    assert r["rel_l2"] < tol, (name, r)
    if name.startswith("nf_"):
        assert r["kind"] == 3, r
    if name.startswith("vf_"):
        assert r["kind"] == 5, r
    if name.startswith("vr_"):
        assert r["kind"] == 6, r
    if name.startswith("ct_") or name in ("direct_k5_8to2", "direct_2seg_to1_f32"):
        assert r["kind"] == 4, r


@pytest.mark.parametrize("pairs", ["1", "3", "7"])
@pytest.mark.parametrize("name", ["vf_k5_48and96_to48_pn", "vf_ragged_37x45", "vf_k3_64to64_pn"])
def test_vfold_row_ranges_crossing_images(name, pairs, monkeypatch):
    """Few CTA pairs = long contiguous row ranges that cross strip and image boundaries (the running sums restart)."""
    monkeypatch.setenv("MPG_VFOLD_PAIRS", pairs)
    kw = dict(CASES[name])
    for key in ("in_dtype", "out_dtype"):
        if kw.get(key, "bf16") == "bf16":
            kw[key] = "f16"
    r = run_case(**kw)
    assert r["finite"] and r["pad_ok"] and r["kind"] == 5, r
    assert r["rel_l2"] < 8e-4, (name, r)


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("name", sorted(k for k in CASES if k.startswith("vr_") and "pair_only" not in k))
def test_vring_single_cta_and_cta_pairs(name, pair, monkeypatch):
    """Every TMEM-ring case in both launch forms: one CTA per strip, and cta_group::2 pairs on adjacent strips with half of
    every weight tile per CTA (odd strip counts leave the second CTA of the last pair without pixels)."""
    monkeypatch.setenv("MPG_VRING_PAIR", pair)
    kw = dict(CASES[name])
    for key in ("in_dtype", "out_dtype"):
        if kw.get(key, "bf16") == "bf16":
            kw[key] = "f16"
    r = run_case(**kw)
    assert r["finite"] and r["pad_ok"] and r["kind"] == 6, r
    assert r["rel_l2"] < (8e-4 if kw["out_dtype"] != "f32" else 2e-5), (name, r)


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("ctas", ["1", "3", "7"])
@pytest.mark.parametrize("name", ["vr_k5_48and96_to48_pn", "vr_ragged_37x45", "vr_k3_64to64_pn", "vr_k5_24to12_pn"])
def test_vring_row_ranges_crossing_images(name, ctas, pair, monkeypatch):
    """Few CTAs = long contiguous row ranges: the TMEM ring wraps many times and crosses strip / image boundaries (the
    partial sums a range leaves in the ring land in slots whose outputs the next range discards)."""
    monkeypatch.setenv("MPG_VRING_CTAS", ctas)
    monkeypatch.setenv("MPG_VRING_PAIR", pair)
    kw = dict(CASES[name])
    for key in ("in_dtype", "out_dtype"):
        if kw.get(key, "bf16") == "bf16":
            kw[key] = "f16"
    r = run_case(**kw)
    assert r["finite"] and r["pad_ok"] and r["kind"] == 6, r
    assert r["rel_l2"] < 8e-4, (name, r)


@pytest.mark.parametrize("pair", ["0", "1"])
def test_vring_is_bit_identical_for_any_row_partition(pair, monkeypatch):
    """Ring positions are a function of the image row, so the order an output's fp32 partial sums are added in does not
    depend on how the rows are cut into per-CTA ranges: the same slices give the same bits whatever the CTA count or the
    number of slices per launch (what keeps a slice-sharded N-GPU volume identical to the single-GPU one)."""
    import numpy as np
    import torch
    from mpgan_b200 import capi
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    n, h, w, cin, cout = 6, 70, 300, 48, 48
    x = torch.randn(n, h, w, cin, generator=g).to(torch.float16).to(dev)
    xs = torch.randn(n, h, w, 96, generator=g).to(torch.float16).to(dev)
    w5 = (torch.randn(5, 5, cin, cout, generator=g) * (np.sqrt(2.0) / np.sqrt(25 * cin))).numpy()
    w1 = (torch.randn(1, 1, 96, cout, generator=g) * (np.sqrt(2.0) / np.sqrt(96))).numpy()
    hd = capi.default_handle(0)
    st = torch.cuda.current_stream().cuda_stream
    monkeypatch.setenv("MPG_VRING_PAIR", pair)

    def run(ctas, n_run):
        monkeypatch.setenv("MPG_VRING_CTAS", ctas)
        plan = capi.ConvPlan(hd, n_run, h, w, [w5, w1], [cin, 96], cout, cout, act="relu", pixel_norm=True, in_dtype=capi.F16,
                             out_dtype=capi.F16, force_kind=6)
        assert plan.kind == capi.KIND_VRING
        y = torch.empty(n_run, h, w, cout, dtype=torch.float16, device=dev)
        plan.run(x[:n_run].contiguous(), xs[:n_run].contiguous(), y, st)
        torch.cuda.synchronize()
        plan.close()
        return y

    ref = run("148", n)
    for ctas in ("1", "5", "37"):
        assert torch.equal(run(ctas, n), ref), ctas
    assert torch.equal(run("148", 2), ref[:2])  # fewer slices per launch: other ranges, same bits
    assert torch.equal(run("3", 1), ref[:1])


def test_row_streaming_auto_rule_picks_wide_images():
    """Auto kind: the row-streaming kernel takes medium / narrow Cout layers on wide images, nothing else."""
    r = run_case(n=4, h=256, w=512, cins=[128], ks=[5], cout=32, act="relu", in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 6 and r["rel_l2"] < 8e-4, r  # weights too large for one CTA: the ring with CTA pairs
    r = run_case(n=4, h=256, w=512, cins=[48], ks=[5], cout=48, act="relu", pixel_norm=True, in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 6 and r["rel_l2"] < 8e-4, r  # the TMEM-ring variant
    r = run_case(n=4, h=256, w=512, cins=[32], ks=[5], cout=8, act="relu", in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 3, r


def test_vfold_is_the_fallback_when_the_ring_is_off(monkeypatch):
    monkeypatch.setenv("MPG_CONV_VRING", "0")
    r = run_case(n=4, h=256, w=512, cins=[128], ks=[5], cout=32, act="relu", in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 5 and r["rel_l2"] < 8e-4, r
    r = run_case(n=1, h=64, w=64, cins=[48], ks=[5], cout=48, act="relu", in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 1, r
    r = run_case(n=1, h=128, w=512, cins=[128], ks=[5], cout=128, act="relu", in_dtype="f16", out_dtype="f16")
    assert r["kind"] == 1, r


def test_flagship_shape_one_slice():
    """ru2_B + shortcut at 512x512 (config 2 resolution), one slice."""
    r = run_case(n=1, h=512, w=512, cins=[128, 32], ks=[5, 1], cout=128, act="relu")
    assert r["kind"] == 1 and r["rel_l2"] < 6e-3, r


@pytest.mark.parametrize("name", ["nfp_k5_128to32_ragged", "nf_k3_64to32_s", "2seg_32k5_128k1_to8"])
def test_nfold_cp_async_producer_variant(name, monkeypatch):
    """The cp.async (LDGSTS) window producer of the tap-folded kernel (off by default: slower than tiled TMA on B200, see
    conv_plan.cu) gives the same results: swizzled destinations, zero fill = SAME padding, CTA-pair relay."""
    monkeypatch.setenv("MPG_NFOLD_CPASYNC", "1")
    kw = dict(CASES[name])
    for key in ("in_dtype", "out_dtype"):
        if kw.get(key, "bf16") == "bf16":
            kw[key] = "f16"
    r = run_case(**kw)
    assert r["finite"] and r["pad_ok"] and r["kind"] == 3, r
    assert r["rel_l2"] < 8e-4, (name, r)


def test_side_output_and_residual_equal_the_two_segment_form():
    """Cross-launch shortcut fusion (mpg_conv_plan_set_side / mpg_conv_plan_run_ex): a 128-channel tcgen05 conv also emits
    y_side = y x W_s (fp32, 8 channels) from its epilogue, and the next block's tap-folded conv adds it as a residual instead
    of re-reading the 128-channel tensor through a 1x1 shortcut segment. Checked against fp64 math on the same operands and
    against the unfused two-segment plan (the results agree to the rounding of the 16-bit intermediate)."""
    import numpy as np
    import torch
    from mpgan_b200 import capi
    from convref import ref_conv
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    n, h, w = 2, 48, 80
    hd = capi.default_handle(0)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(n, h, w, 128, generator=g).to(torch.float16).to(dev)
    xs = torch.randn(n, h, w, 32, generator=g).to(torch.float16).to(dev)
    wp = (torch.randn(5, 5, 128, 128, generator=g) * (np.sqrt(2.0) / np.sqrt(3200))).numpy()
    wps = (torch.randn(1, 1, 32, 128, generator=g) * (np.sqrt(2.0) / np.sqrt(32))).numpy()
    shp = (torch.randn(128, generator=g) * 0.1).numpy()
    ws_next = (torch.randn(1, 1, 128, 8, generator=g) * (np.sqrt(2.0) / np.sqrt(128))).numpy()
    producer = capi.ConvPlan(hd, n, h, w, [wp, wps], [128, 32], 128, 128, act="relu", shift=shp, in_dtype=capi.F16, out_dtype=capi.F16)
    assert producer.kind == capi.KIND_TCGEN05
    producer.set_side(ws_next[0, 0])
    y = torch.empty(n, h, w, 128, dtype=torch.float16, device=dev)
    side = torch.full((n, h, w, 8), float("nan"), dtype=torch.float32, device=dev)
    producer.run_ex(x, xs, y, y_side=side, stream=st)
    with pytest.raises(capi.MpgError):
        producer.run(x, xs, y, st)  # a plan with a side output must be given the side tensor
    torch.cuda.synchronize()
    yref = ref_conv([x.float(), xs.float()], [wp, wps], [None, None], shp, "relu", False, 1, round_w=torch.float16)
    assert float(torch.linalg.norm(y.double() - yref) / torch.linalg.norm(yref)) < 8e-4
    side_ref = torch.einsum("nhwc,ck->nhwk", yref, torch.from_numpy(ws_next[0, 0]).double().to(dev))
    rel = float(torch.linalg.norm(side.double() - side_ref) / torch.linalg.norm(side_ref))
    assert torch.isfinite(side).all() and rel < 2e-5, rel  # fp32 values straight from the accumulators
    # consumer: 5x5 32->8 (+ the 1x1 128->8 shortcut): residual form vs two-segment form
    x32 = torch.randn(n, h, w, 32, generator=g).to(torch.float16).to(dev)
    w5 = (torch.randn(5, 5, 32, 8, generator=g) * (np.sqrt(2.0) / np.sqrt(800))).numpy()
    shc = (torch.randn(8, generator=g) * 0.1).numpy()
    fused = capi.ConvPlan(hd, n, h, w, [w5], [32], 8, 8, act="relu", shift=shc, in_dtype=capi.F16, out_dtype=capi.F16, force_kind=3)
    assert fused.kind == capi.KIND_NFOLD
    o1 = torch.empty(n, h, w, 8, dtype=torch.float16, device=dev)
    fused.run_ex(x32, None, o1, residual=side, stream=st)
    fused_vf = capi.ConvPlan(hd, n, h, w, [w5], [32], 8, 8, act="relu", shift=shc, in_dtype=capi.F16, out_dtype=capi.F16, force_kind=5)
    assert fused_vf.kind == capi.KIND_VFOLD
    o3 = torch.empty(n, h, w, 8, dtype=torch.float16, device=dev)
    fused_vf.run_ex(x32, None, o3, residual=side, stream=st)
    two = capi.ConvPlan(hd, n, h, w, [w5, ws_next], [32, 128], 8, 8, act="relu", shift=shc, in_dtype=capi.F16, out_dtype=capi.F16)
    o2 = torch.empty(n, h, w, 8, dtype=torch.float16, device=dev)
    two.run(x32, y, o2, st)
    torch.cuda.synchronize()
    want = torch.relu(ref_conv([x32.float()], [w5], [None], shc, None, False, 1, round_w=torch.float16) + side_ref)
    for got in (o1, o2, o3):
        assert float(torch.linalg.norm(got.double() - want) / torch.linalg.norm(want)) < 1.5e-3
    for pl in (producer, fused, two, fused_vf):
        pl.close()
    # plans that cannot carry a side output say so
    small = capi.ConvPlan(hd, 1, 32, 32, [np.zeros((3, 3, 64, 64), np.float32)], [64], 64, 64, in_dtype=capi.F16, out_dtype=capi.F16)
    with pytest.raises(capi.MpgError):
        small.set_side(np.zeros((64, 8), np.float32))
    small.close()
