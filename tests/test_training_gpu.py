"""GPU parity of the 4x training step (generator + spatial discriminator fwd/bwd, losses, TF1 Adam, BN moving
averages) against the fp64 autograd oracle (oracle/training.py), through the C ABI (mpg_train_*)."""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi, training as T
from oracle import networks as on, training as ot

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("bn", [True, False])
def test_training_iteration_matches_oracle(bn):
    L, u, B = 8, 4, 4
    S = L * u
    rng = np.random.default_rng(5)
    hp = dict(kk=5.0, kk2=1e-5, seed=9, weight_dld=1.0)
    batches = [(rng.random((B, L * L * 4), dtype=np.float32), rng.random((B, S * S), dtype=np.float32)) for _ in range(2)]
    tr = T.Trainer4x(L, u, B, seed=9, batch_norm=bn)
    values = {k: v.copy() for k, v in tr.values().items()}
    cfg = on.make_cfg_4x(L, upRes=u, upsampling_mode=2, batch_norm=bn)
    od, og_ = ot.Adam(2e-4, 0.5), ot.Adam(2e-4, 0.5)
    for it in range(2):  # two iterations: the second one exercises non-zero Adam moments and moved BN statistics
        ref = ot.train_iteration(values, [batches[it]], [batches[it]], cfg, hp, od, og_)
        got = tr.iteration([batches[it]], [batches[it]], kk=hp["kk"], kk2=hp["kk2"])
        assert abs(got["disc_loss"] - ref["disc_loss"]) <= 1e-4 * max(1.0, abs(ref["disc_loss"])), (got, ref)
        assert abs(got["gen_loss"] - ref["gen_loss"]) <= 1e-4, (got, ref)
        assert abs(got["gen_loss_complete"] - ref["gen_loss_complete"]) <= 2e-4 * max(1.0, abs(ref["gen_loss_complete"])), (got, ref)
        gg = tr.grads("g")
        for name, g in ref["grads_g"].items():
            scale = np.abs(g).max()
            if scale < 1e-12:
                continue
            assert _rel(gg[name], g) <= 2e-3, (it, name, _rel(gg[name], g))
        now = tr.values()
        lr = 2e-4
        ref_g = dict(ref["grads_d"])
        ref_g.update(ref["grads_g"])
        for name, v in values.items():
            d = np.abs(now[name].astype(np.float64) - np.asarray(v, np.float64))
            if name.endswith(("moving_mean", "moving_variance")):
                assert d.max() <= 1e-5 * max(1.0, np.abs(v).max()), (it, name, d.max())
            else:
                # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is ~0 may differ by 2 lr
                assert d.max() <= 2.2 * lr * (it + 1), (it, name, d.max())
                # a bias in front of a batch norm has a mathematically zero gradient: fp32 rounding noise of the size
                # of Adam's epsilon makes it random-walk by < lr per step (TF fp32 does the same) - only bounded above
                if np.abs(ref_g[name]).max() > 1e-9:
                    assert (d > 2e-6).sum() <= max(2, 0.02 * d.size), (it, name, int((d > 2e-6).sum()), d.size)


def test_adam_kernel_tf1_form():
    h = capi.default_handle(0)
    n = 1000
    rng = np.random.default_rng(0)
    p, g = rng.standard_normal(n).astype(np.float32), (rng.standard_normal(n) * 1e-3).astype(np.float32)
    m, v = rng.random(n).astype(np.float32) * 1e-3, rng.random(n).astype(np.float32) * 1e-6
    dp, dg, dm, dv = [torch.from_numpy(a.copy()).cuda() for a in (p, g, m, v)]
    capi.train_call("adam", h, dp, dg, dm, dv, n, 1.5e-4, 0.5, 0.999, 1e-8, 0)
    m2 = 0.5 * m.astype(np.float64) + 0.5 * g
    v2 = 0.999 * v.astype(np.float64) + 0.001 * g.astype(np.float64) ** 2
    ref = p - 1.5e-4 * m2 / (np.sqrt(v2) + 1e-8)
    np.testing.assert_allclose(dp.cpu().numpy(), ref, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(dm.cpu().numpy(), m2, rtol=1e-5)


def test_train_conv_kernels_vs_torch_autograd():
    """fwd / dgrad / wgrad of every conv geometry of the training graphs vs torch autograd (fp64 on CPU)."""
    h = capi.default_handle(0)
    from oracle import tf_ops
    rng = np.random.default_rng(1)
    for (n, hh, cin, cout, k, s, up) in [(2, 16, 4, 8, 5, 1, 4), (2, 16, 8, 32, 5, 1, 1), (1, 16, 32, 128, 1, 1, 1),
                                          (2, 16, 2, 32, 4, 2, 1), (2, 8, 64, 128, 4, 2, 1), (2, 8, 128, 256, 4, 1, 1),
                                          (2, 12, 2, 1, 5, 1, 1)]:
        sh = hh // up
        x = rng.standard_normal((n, sh, sh, cin)).astype(np.float32)
        w = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
        xu = tf_ops.resize_nearest(xt, hh, hh) if up > 1 else xt
        yt = tf_ops.conv2d_same(xu, wt, s) + torch.tensor(b, dtype=torch.float64)
        dy = rng.standard_normal(tuple(yt.shape)).astype(np.float32)
        yt.backward(torch.tensor(dy, dtype=torch.float64))
        oh = yt.shape[1]
        dx_, dw_, dy_d = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(dy).cuda()
        y_d = torch.empty((n, oh, oh, cout), device="cuda")
        capi.train_call("conv_fwd", h, dx_, dw_, torch.from_numpy(b).cuda(), y_d, n, hh, hh, cin, cout, k, s, up, 0)
        assert _rel(y_d.cpu().numpy(), yt.detach().numpy()) < 1e-5, ("fwd", cin, cout, k, s)
        gw = torch.zeros_like(dw_)
        gb = torch.zeros(cout, device="cuda")
        scratch = torch.zeros(1024, dtype=torch.float64, device="cuda")
        capi.train_call("conv_wgrad", h, dx_, dy_d, gw, gb, scratch, n, hh, hh, cin, cout, k, s, up, 0)
        assert _rel(gw.cpu().numpy(), wt.grad.numpy()) < 1e-5, ("wgrad", cin, cout, k, s)
        assert _rel(gb.cpu().numpy(), dy.astype(np.float64).sum(axis=(0, 1, 2))) < 1e-5
        if up == 1:
            gx = torch.empty_like(dx_)
            capi.train_call("conv_dgrad", h, dy_d, dw_, gx, n, hh, hh, cin, cout, k, s, 0, 0)
            assert _rel(gx.cpu().numpy(), xt.grad.numpy()) < 1e-5, ("dgrad", cin, cout, k, s)


def test_training_tensor_core_mode_tracks_oracle():
    """precision="fp16": wide convs forward/dgrad on tcgen05 (fp16 activations, bf16 gradients). One iteration must
    stay close to the fp64 oracle (losses 1e-2 relative, generator gradients of the wide layers 5e-2 rel-L2)."""
    L, u, B = 8, 4, 4
    S = L * u
    rng = np.random.default_rng(6)
    hp = dict(kk=5.0, kk2=1e-5, seed=11, weight_dld=1.0)
    xb, yb = rng.random((B, L * L * 4), dtype=np.float32), rng.random((B, S * S), dtype=np.float32)
    tr = T.Trainer4x(L, u, B, seed=11, precision="fp16")
    assert any(c.fast for rb in tr.rbs for c in rb)
    values = {k: v.copy() for k, v in tr.values().items()}
    cfg = on.make_cfg_4x(L, upRes=u, upsampling_mode=2, batch_norm=True)
    ref = ot.train_iteration(values, [(xb, yb)], [(xb, yb)], cfg, hp, ot.Adam(2e-4, 0.5), ot.Adam(2e-4, 0.5))
    got = tr.iteration([(xb, yb)], [(xb, yb)], kk=hp["kk"], kk2=hp["kk2"])
    assert abs(got["disc_loss"] - ref["disc_loss"]) <= 1e-2 * max(1.0, abs(ref["disc_loss"])), (got, ref)
    assert abs(got["gen_loss_complete"] - ref["gen_loss_complete"]) <= 1e-2 * max(1.0, abs(ref["gen_loss_complete"])), (got, ref)
    gg = tr.grads("g")
    for name in ("generator/g_cB1/weight", "generator/g_cA1/weight", "generator/g_cA2/weight", "generator/g_cA0/weight"):
        # measured 3-5 % on g_cB1 (bf16 operands in dgrad and wgrad; the split-K atomics make the last digits jitter)
        assert _rel(gg[name], ref["grads_g"][name]) <= 8e-2, (name, _rel(gg[name], ref["grads_g"][name]))


@pytest.mark.parametrize("shape", [(2, 12, 64, 128, 128, 5), (3, 9, 32, 32, 128, 5), (2, 16, 16, 128, 32, 5),
                                   (2, 8, 32, 32, 128, 1), (1, 10, 48, 64, 128, 3), (2, 7, 32, 128, 64, 3)])
def test_wgrad_tensor_core_vs_torch_autograd(shape):
    """mpg_train_conv_wgrad_tc (tcgen05, MN-major operands straight from NHWC) vs torch autograd in fp64 on the SAME
    bf16-rounded inputs: the only difference left is fp32 accumulation order, so the bound is tight (2e-5 rel-L2).
    Covers both operand roles (Cin = 128 / Cout = 128), the 64-byte-swizzle 32-channel operand, k = 1/3/5, non-square
    images and a second call accumulating on top of the first (dw += ...)."""
    n, hh, ww, cin, cout, k = shape
    h = capi.default_handle(0)
    from oracle import tf_ops
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((n, hh, ww, cin)).astype(np.float32)).to(torch.bfloat16)
    dy = torch.from_numpy((rng.standard_normal((n, hh, ww, cout)) * 1e-3).astype(np.float32)).to(torch.bfloat16)
    wt = torch.zeros((k, k, cin, cout), dtype=torch.float64, requires_grad=True)
    yt = tf_ops.conv2d_same(x.double(), wt, 1)
    yt.backward(dy.double())
    ref = wt.grad.numpy()
    gw = torch.zeros((k, k, cin, cout), device="cuda")
    xd, dyd = x.cuda().contiguous(), dy.cuda().contiguous()
    capi.train_call("conv_wgrad_tc", h, xd, dyd, gw, n, hh, ww, cin, cout, k, 0)
    torch.cuda.synchronize()
    assert _rel(gw.cpu().numpy(), ref) < 2e-5, (shape, _rel(gw.cpu().numpy(), ref))
    capi.train_call("conv_wgrad_tc", h, xd, dyd, gw, n, hh, ww, cin, cout, k, 0)
    torch.cuda.synchronize()
    assert _rel(gw.cpu().numpy(), 2.0 * ref) < 2e-5, shape


def test_wgrad_tensor_core_rejects_unsupported_shapes():
    h = capi.default_handle(0)
    x = torch.zeros((1, 8, 16, 32), dtype=torch.bfloat16, device="cuda")
    dy = torch.zeros((1, 8, 16, 8), dtype=torch.bfloat16, device="cuda")
    gw = torch.zeros((5, 5, 32, 8), device="cuda")
    with pytest.raises(Exception):
        capi.train_call("conv_wgrad_tc", h, x, dy, gw, 1, 8, 16, 32, 8, 5, 0)


def test_training_graph_replay_matches_eager():
    """graphs=True: step 1 runs eagerly, step 2 is captured, steps 3+ are replays of the captured CUDA graph (Adam step
    size read from a device scalar). Four iterations must track an eager trainer (only the atomics' summation order
    differs)."""
    L, u, B = 8, 4, 4
    S = L * u
    rng = np.random.default_rng(8)
    batches = [(rng.random((B, L * L * 4), dtype=np.float32), rng.random((B, S * S), dtype=np.float32)) for _ in range(4)]
    te = T.Trainer4x(L, u, B, seed=3)
    tg = T.Trainer4x(L, u, B, seed=3, graphs=True)
    assert tg.use_graphs
    for it in range(4):
        a = te.iteration([batches[it]], [batches[it]])
        b = tg.iteration([batches[it]], [batches[it]])
        for k in a:
            assert abs(a[k] - b[k]) <= 3e-4 * max(1.0, abs(a[k])), (it, k, a[k], b[k])
    va, vb = te.values(), tg.values()
    for name in va:
        # biases in front of a batch norm have an exactly-zero gradient; Adam turns their rounding noise into +-lr steps,
        # so only the filter tensors are compared
        if name.endswith("/weight"):
            assert _rel(vb[name], va[name]) <= 3e-3, (name, _rel(vb[name], va[name]))
    # (graph keys carry which weight sets the captured body re-packs: the steady-state D and G bodies are captured)
    assert any(isinstance(v, dict) for k, v in tg._graphs.items() if k[0] == "d")
    assert any(isinstance(v, dict) for k, v in tg._graphs.items() if k[0] == "g")


@pytest.mark.parametrize("stride,cin,cout", [(2, 2, 32), (2, 32, 64), (1, 128, 256)])
def test_k4_conv_through_embedded_5x5_tensor_core_plan(stride, cin, cout):
    """disc_binclass convs (k = 4, stride 2 / 1, GAN/multipassGAN-4x.py:593-614) on the tcgen05 kernel: TF's SAME window of a
    4x4 kernel = taps -1..+2 of a 5x5 one (mpg_conv_plan_update_ex embeds the weights), a stride-2 layer = the stride-1 plan
    sampled at even positions (mpg_train_pick), its input gradient = the plan's dgrad of dy scattered onto those positions
    (mpg_train_stuff16). Against torch conv2d / its autograd on the same 16-bit-rounded operands."""
    import torch.nn.functional as F
    from mpgan_b200 import capi
    from convref import tf_same_pad
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    n, h, w = 3, 16, 16
    hd = capi.default_handle(0)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(n, h, w, cin, generator=g)
    wt = (torch.randn(4, 4, cin, cout, generator=g) * (np.sqrt(2.0) / np.sqrt(16 * cin))).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
    cpi, cpo = -(-cin // 8) * 8, -(-cout // 8) * 8
    x16 = torch.zeros(n, h, w, cpi, dtype=torch.float16, device=dev)
    x16[..., :cin] = x.to(torch.float16).to(dev)
    oh, ow = -(-h // stride), -(-w // stride)
    lin = torch.full((n, oh, ow, cout), float("nan"), device=dev)
    for c0 in range(0, cout, 128):
        cc = min(128, cout - c0)
        pl = capi.ConvPlan(hd, n, h, w, [np.zeros((5, 5, cin, cc), np.float32)], [cpi], cc, cc, act=None,
                           shift=np.zeros(cc, np.float32), in_dtype=capi.F16, out_dtype=capi.F32, force_kind=1)
        pl.update(wt, mode0=0, shift=bias[c0:c0 + cc], stream=st, src_k=4, src_cout=cout, cout_off=c0)
        full = torch.empty(n, h, w, cc, device=dev)
        pl.run(x16, None, full, st)
        capi.train_call("pick", hd, full, lin, n, oh, ow, cc, stride, cc, cout, c0, st)
        pl.close()
    torch.cuda.synchronize()
    xr = x16[..., :cin].double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.to(torch.float16).double().permute(3, 2, 0, 1)
    pt, pb = tf_same_pad(h, 4, stride)
    plft, prt = tf_same_pad(w, 4, stride)
    ref = F.conv2d(F.pad(xr, (plft, prt, pt, pb)), wr, bias.double(), stride=stride)
    assert float(torch.linalg.norm(lin.double() - ref.permute(0, 2, 3, 1)) / torch.linalg.norm(ref)) < 1e-3
    # input gradient
    dy = torch.randn(n, oh, ow, cout, generator=g).to(dev)
    g16 = torch.empty(n, h, w, cpo, dtype=torch.bfloat16, device=dev)
    capi.train_call("stuff16", hd, dy, g16, capi.BF16, n, oh, ow, cout, stride, cpo, st)
    pd = capi.ConvPlan(hd, n, h, w, [np.zeros((5, 5, cout, cin), np.float32)], [cpo], cin, cin, act=None,
                       shift=np.zeros(cin, np.float32), in_dtype=capi.BF16, out_dtype=capi.F32, force_kind=1)
    pd.update(wt, mode0=1, stream=st, src_k=4)
    dx = torch.empty(n, h, w, cin, device=dev)
    pd.run(g16, None, dx, st)
    torch.cuda.synchronize()
    pd.close()
    dyr = dy.to(torch.bfloat16).double().permute(0, 3, 1, 2)
    wr_b = wt.to(torch.bfloat16).double().permute(3, 2, 0, 1)
    ref2 = F.conv2d(F.pad(xr, (plft, prt, pt, pb)), wr_b, None, stride=stride)
    (gx,) = torch.autograd.grad(ref2, xr, dyr)
    assert float(torch.linalg.norm(dx.double() - gx.permute(0, 2, 3, 1)) / torch.linalg.norm(gx)) < 2e-3


def test_fast_mode_tracks_the_fp32_mode_over_many_iterations():
    """Drift of precision="fp16" (tensor-core forward / dgrad / wgrad, tensor-core discriminator convs) against the fp32
    parity mode over 12 loop bodies on fresh batches: every loss stays within 5e-3 (measured 1e-3) and the trained filters within 5e-3
    rel-L2 of the fp32 run (one-iteration errors do not compound)."""
    L, u, B = 8, 4, 4
    S = L * u
    rng = np.random.default_rng(21)
    batches = [(rng.random((B, L * L * 4), dtype=np.float32), rng.random((B, S * S), dtype=np.float32)) for _ in range(12)]
    t32 = T.Trainer4x(L, u, B, seed=5, precision="fp32")
    t16 = T.Trainer4x(L, u, B, seed=5, precision="fp16")
    worst = 0.0
    for xb, yb in batches:
        a = t32.iteration([(xb, yb)], [(xb, yb)])
        b = t16.iteration([(xb, yb)], [(xb, yb)])
        for k in ("disc_loss", "gen_loss_complete", "gen_l1_loss_scaled"):
            d = abs(a[k] - b[k]) / max(1.0, abs(a[k]))
            worst = max(worst, d)
            assert d <= 5e-3, (k, a[k], b[k])  # measured 1.0e-3
    va, vb = t32.values(), t16.values()
    for name in va:
        if name.endswith("/weight"):
            assert _rel(vb[name], va[name]) <= 5e-3, (name, _rel(vb[name], va[name]))
    print("worst relative loss deviation over 12 iterations: %.3e" % worst)
