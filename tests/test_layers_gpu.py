"""GPU parity of the library-level layer API through the engine and the C ABI:
  * GAN.residual_block as written in tools_wscale/GAN.py:126-147 (tanh default, BN on/off) -- the scripts use their
    own resBlock copies, this is the library version
  * GAN.avg_depool (tools_wscale/GAN.py:528-552) with all three tf.image.ResizeMethod values: 0 bilinear, 1 nearest,
    2 bicubic on a multi-channel tensor that feeds further convolutions
against the fp64 oracle restatement (oracle/gan.py, oracle/tf_ops.py).
"""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi, engine, graph as G, weights as W
from mpgan_b200.GAN import GAN, lrelu, relu
from oracle import gan as og
from oracle import tf_ops
from oracle_nets import err_stats

pytestmark = pytest.mark.gpu

TOL = {"fp32": (1e-4, 1e-4), "fp16": (5e-3, 2e-2)}


def _check(name, got, ref, precision):
    st = err_stats(got, ref)
    rel_tol, abs_tol = TOL[precision]
    print("%s [%s] rel_l2=%.3e max_abs=%.3e ref_max=%.3f" % (name, precision, st["rel_l2"], st["max_abs"], st["ref_max"]))
    assert np.isfinite(got).all()
    assert st["rel_l2"] <= rel_tol, (name, precision, st)
    assert st["max_abs"] <= abs_tol * max(1.0, st["ref_max"]), (name, precision, st)


def _build_rb(make_gan, img, act, bn):
    """Two chained library residual blocks; the first one uses the default activation when act is None."""
    gan = make_gan(img)
    kw = {} if act is None else dict(activation_function=act)
    gan.residual_block(8, 16, [3, 3], name="RB1", batch_norm=bn, train=False, **kw)
    out, lin = gan.residual_block(16, 8, [5, 5], name="RB2", batch_norm=bn, train=False, **kw)
    return gan, out, lin


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("bn", [False, True])
@pytest.mark.parametrize("act", [None, "relu", "lrelu"])
def test_library_residual_block(precision, bn, act):
    B, H, Wd, C = 3, 24, 40, 3
    G.reset_default_graph()
    x = G.placeholder([None, H * Wd * C], "x")
    acts = {None: None, "relu": relu, "lrelu": lrelu}
    gan, out, lin = _build_rb(lambda im: GAN(im), G.reshape(x, [-1, H, Wd, C]), acts[act], bn)
    assert gan.layer is out and lin.node.op == "add"
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), 11), 11)
    rng = np.random.default_rng(4)
    xv = (rng.standard_normal((B, H * Wd * C)) * 0.5).astype(np.float32)
    net = engine.CompiledNet(out, w, B, precision=precision)
    got = net.run({"x": torch.from_numpy(xv).cuda()}).float().cpu().numpy().reshape(B, H, Wd, -1)[..., :8]
    net.close()

    oacts = {None: None, "relu": og.relu, "lrelu": og.lrelu}
    ctx = og.Context(og.VarStore(values=w), torch.float64)
    _, oref, _ = _build_rb(lambda im: og.GAN(im, ctx), torch.from_numpy(xv).double().reshape(B, H, Wd, C), oacts[act], bn)
    _check("residual_block act=%s bn=%s" % (act or "tanh(default)", bn), got, oref.numpy(), precision)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_avg_depool_modes(precision, mode):
    """conv -> avg_depool(mode, scale 2) -> conv: the resize of a 16-channel tensor feeding another convolution."""
    B, H, Wd, C = 2, 12, 20, 4
    G.reset_default_graph()
    x = G.placeholder([None, H * Wd * C], "x")

    def build(gan, a):
        gan.convolutional_layer(16, [3, 3], a, name="c1")
        up = gan.avg_depool(mode=mode, scale=[2])
        y, _ = gan.convolutional_layer(8, [3, 3], a, name="c2")
        return up, y

    up, y = build(GAN(G.reshape(x, [-1, H, Wd, C])), relu)
    assert up.shape == (None, 2 * H, 2 * Wd, 16)
    w = W.init_graph_variables(G.get_default_graph(), 12)
    rng = np.random.default_rng(5)
    xv = rng.standard_normal((B, H * Wd * C)).astype(np.float32)
    net = engine.CompiledNet(y, w, B, precision=precision)
    got = net.run({"x": torch.from_numpy(xv).cuda()}).float().cpu().numpy().reshape(B, 2 * H, 2 * Wd, -1)[..., :8]
    net.close()
    ctx = og.Context(og.VarStore(values=w), torch.float64)
    _, oy = build(og.GAN(torch.from_numpy(xv).double().reshape(B, H, Wd, C), ctx), og.relu)
    _check("avg_depool mode %d" % mode, got, oy.numpy(), precision)


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("factor", [(2, 2), (4, 3)])
def test_resize_images_kernel(mode, factor):
    """mpg_resize_images alone (fp32 in / out) against oracle.tf_ops (TF1 legacy bilinear / bicubic rules, App. B.6)."""
    n, h, w_, c = 2, 9, 7, 5
    oh, ow = h * factor[0], w_ * factor[1]
    rng = np.random.default_rng(6)
    x = rng.standard_normal((n, h, w_, c)).astype(np.float32)
    hd = capi.default_handle(0)
    xd = torch.from_numpy(x).cuda()
    out = torch.full((n, oh, ow, 8), 7.0, dtype=torch.float32, device="cuda")
    plan = capi.BicubicPlan(hd, h, w_, oh, ow) if mode == 2 else None
    capi.resize_images(hd, xd, capi.F32, c, c, n, h, w_, out, capi.F32, 8, oh, ow, mode, plan,
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    fn = tf_ops.resize_bilinear_tf1 if mode == 0 else tf_ops.resize_bicubic_tf1
    ref = fn(torch.from_numpy(x).double(), oh, ow).numpy()
    assert np.abs(got[..., :c] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    assert (got[..., c:] == 0).all()  # padded channels are written as zeros


@pytest.mark.parametrize("mode", [-1, 0, 2])
@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
def test_dens_out_fused_density_output(mode, dt):
    """mpg_dens_out = g_cdensOut (1x1 conv to one channel, GAN/multipassGAN-out.py:282) + the additive residual (:327-332)
    in one launch, against the separate oracle ops."""
    import numpy as np
    from mpgan_b200 import capi
    from oracle import tf_ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    n, h, w, cin, cs = 2, 40, 48, 12, 16
    tdt = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[dt]
    code = {"f16": capi.F16, "bf16": capi.BF16, "f32": capi.F32}[dt]
    x = torch.zeros(n, h, w, cs)
    x[..., :cin] = torch.randn(n, h, w, cin, generator=g)
    xd = x.to(tdt).to(dev)
    wv = torch.randn(cin, generator=g).numpy().astype(np.float32)
    bias = 0.1
    want = torch.einsum("nhwc,c->nhw", xd.double()[..., :cin].cpu(), torch.from_numpy(wv).double()) + bias
    hd = capi.default_handle(0)
    plan, src, sh, sw = None, None, h, w
    if mode == 0:
        src = torch.randn(n, h, w, 5, generator=g).to(dev)
        want = want + src.double().cpu()[..., 2]
    elif mode == 2:
        sh, sw = h // 8, w // 8
        src = torch.randn(n, sh, sw, 5, generator=g).to(dev)
        up = tf_ops.resize_bicubic_tf1(src.cpu().double()[..., 2:3], h, w)
        want = want + up[..., 0]
        plan = capi.BicubicPlan(hd, sh, sw, h, w)
    out = torch.full((n, h, w), float("nan"), device=dev)
    capi.dens_out(hd, xd, code, cs, cin, wv, bias, src, capi.F32, 5, 2, mode, plan, n, h, w, sh, sw, out,
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = float((out.double().cpu() - want).abs().max())
    assert err < 2e-5 * max(1.0, float(want.abs().max())), err
