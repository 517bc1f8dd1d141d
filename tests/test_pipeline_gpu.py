"""GPU parity of the whole multi-pass volume pipelines vs the line-by-line numpy oracle."""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import pipeline as P, synth
from oracle import pipeline as op
from oracle_nets import err_stats, oracle_gen_resnet, oracle_growing_gen

pytestmark = pytest.mark.gpu

# north_star: 16-bit path rel-L2 <= 5e-3, max-abs <= 2e-2 (x output scale: random-init outputs exceed [0,1]);
# fp32 path <= 1e-4.  bf16 operands do NOT meet 5e-3 on these 8..31-layer chains (measured 0.4-2.5 %): it is
# kept as an opt-in mode with a loose regression bound; the default 16-bit type is IEEE fp16.
TOL = {"fp32": (1e-4, 1e-4), "fp16": (5e-3, 2e-2), "bf16": (5e-2, 2e-1)}


def _check(name, got, ref, precision, thr=0.0005):
    """`thr`: the output threshold (v < 0.0005 -> 0) is a discontinuity, so an element within rounding of it
    may flip; that adds at most `thr` to the max-abs bound."""
    st = err_stats(got, ref)
    rel_tol, abs_tol = TOL[precision]
    print("%s [%s] rel_l2=%.3e max_abs=%.3e ref_max=%.3f" % (name, precision, st["rel_l2"], st["max_abs"], st["ref_max"]))
    assert np.isfinite(got).all()
    assert st["rel_l2"] <= rel_tol and st["max_abs"] <= abs_tol * max(1.0, st["ref_max"]) + thr, (name, precision, st)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_two_pass_4x_volume(precision):
    L, u = 8, 4
    x = synth.synthetic_volume(L, seed=11)
    w1, w2 = P.make_weights_4x(L, 7, upRes=u, randomize_bn=True)
    mp = P.MultiPass4x(L, w1, w2, upRes=u, precision=precision, batch=8)
    got = mp(x).cpu().numpy()
    ref, p1 = op.two_pass_4x(oracle_gen_resnet(w1, L, 2), oracle_gen_resnet(w2, L, 1), x, u, return_intermediate=True)
    got1 = mp.pass1_only(x).cpu().numpy()
    _check("4x pass 1", got1, p1, precision)
    _check("4x two-pass", got, ref, precision)


def test_two_pass_4x_velscale_quirk_fp32():
    """App. D.10: velScale reaches only vy,vz of the pass-2 velocity array."""
    L, u = 8, 4
    x = synth.synthetic_volume(L, seed=12)
    w1, w2 = P.make_weights_4x(L, 8, upRes=u)
    got = P.MultiPass4x(L, w1, w2, upRes=u, precision="fp32", velScale=1.5)(x).cpu().numpy()
    ref = op.two_pass_4x(oracle_gen_resnet(w1, L, 2), oracle_gen_resnet(w2, L, 1), x, u, velScale=1.5)
    _check("4x velScale", got, ref, "fp32")


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_out_two_nets_8x(precision):
    L, u = 4, 8
    x = np.random.default_rng(13).random((L, L, L, 4), dtype=np.float32)
    w = P.make_weights_out(L, 9, upRes=u, nets=(1, 2))
    mp = P.MultiPassOut(L, w, upRes=u, precision=precision, threshold=0.0)
    got = mp(x).cpu().numpy()
    n1 = oracle_growing_gen(w[1], 1, P.SHIPPED_8X[1], L, upRes=u)
    n2 = oracle_growing_gen(w[2], 2, P.SHIPPED_8X[2], L, upRes=u)
    ref = op.out_generate3d(x, u, n1, n2, None, transposeAxis=0, add_adj_idcs1=True, threshold=False)
    assert np.abs(ref).max() > 0.1
    _check("out.py 8x nets 1+2", got, ref, precision)


@pytest.mark.parametrize("ta", [0, 1])
def test_out_three_nets_axis_bookkeeping_fp32(ta):
    """Small 2x nets, all three passes, two transposeAxis settings: catches any axis/channel mix-up."""
    L, u = 4, 2
    specs = {1: P.NetSpec(True, True, 16, 16, 3, True), 2: P.NetSpec(True, False, 16, 16, 3),
             3: P.NetSpec(False, False, 16, 8, 3)}
    x = np.random.default_rng(14).random((L, L, L, 4), dtype=np.float32)
    w = P.make_weights_out(L, 10, upRes=u, specs=specs, nets=(1, 2, 3))
    mp = P.MultiPassOut(L, w, upRes=u, specs=specs, precision="fp32", transposeAxis=ta, batches=(8, 2, 2))
    got = mp(x).cpu().numpy()
    nets = [oracle_growing_gen(w[i], i, specs[i], L, upRes=u) for i in (1, 2, 3)]
    ref = op.out_generate3d(x, u, nets[0], nets[1], nets[2], transposeAxis=ta, add_adj_idcs1=True)
    _check("out.py 3 nets ta=%d" % ta, got, ref, "fp32")


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_two_pass_4x_tiled_apply_is_bit_identical(precision):
    """SURVEY 8 a19 wired into the pipeline: both passes applied through overlapping tiles (mpg_tiles_cut -> network on
    batch*nt*nt tiles -> mpg_tiles_stitch_overlap, halo = receptive-field radius 16 px) must reproduce the whole-slice
    apply exactly (same kernels, same accumulation order per pixel; frame-edge tiles keep their zero-padded edge)."""
    L, u = 24, 4  # S = 96: pass-1 tiles 8+2*4 low-res px (2x2 per slice), pass-2 tiles 32+2*16 px (2x2 per slice)
    x = synth.synthetic_volume(L, seed=21)
    w1, w2 = P.make_weights_4x(L, 5, upRes=u, randomize_bn=True)
    whole = P.MultiPass4x(L, w1, w2, upRes=u, precision=precision, batch=4)
    ref = whole(x).cpu().numpy().copy()
    tiled = P.MultiPass4x(L, w1, w2, upRes=u, precision=precision, batch=4, tile=(8, 32))
    assert tiled.p1.net.nt == 2 and tiled.p2.net.nt == 2
    got = tiled(x).cpu().numpy()
    assert np.isfinite(got).all() and float(np.abs(ref).max()) > 0
    assert np.array_equal(got, ref), float(np.abs(got - ref).max())
