"""GPU parity of the 8x progressive-growing trainer pieces (multi-pass-gan_b200/training8x.py) against the
fp64 autograd oracle (oracle/training8x.py, itself pinned by executing the reference's growing_disc on the TF1 shim):
critic forward for both architectures and several growing percentages, the WGAN-GP discriminator loss and ALL its parameter
gradients (the gradient penalty differentiates a gradient: tangent-pass formulation vs torch double backward), input
gradients, the staged Adam optimizers and the weight EMA."""
import json
import os

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import training8x as t8
from oracle import gan as og
from oracle import training as ot
from oracle import training8x as o8

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "growdisc.npz"))


def _setup(tag, batch=2):
    c = json.loads(str(GOLD[tag + "_cfg"]))
    mode = c.get("upsampling_mode", 2)
    cfg = o8.Cfg8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], c["filterSize"], c["first_nn_arch"], upsampling_mode=mode)
    store = og.VarStore(seed=c["seed"])
    x = torch.from_numpy(GOLD[tag + "_x"]).double()
    y = torch.from_numpy(GOLD[tag + "_y"]).double()
    o8.growing_disc(y, x, 1.0, og.Context(store, torch.float64), cfg)  # creates every variable
    d = t8.GrowingDisc(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], c["filterSize"], c["first_nn_arch"], batch=batch,
                       values=store.values, upsampling_mode=mode)
    assert {n for n, *_ in d.ps.specs} == set(store.values)
    return c, cfg, store, x, y, d


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("tag", ["gd_first", "gd_plain", "gd_second"])
def test_growing_disc_forward_matches_the_reference_vectors(tag):
    c, cfg, store, x, y, d = _setup(tag)
    dev = d.cx.device
    d.cx.st = torch.cuda.current_stream(dev).cuda_stream
    d.refresh()
    for k, pct in enumerate(c["percentages"]):
        logits, _ = d.forward(x.float().to(dev), y.float().to(dev), pct)
        ref = GOLD["%s_p%d_logits" % (tag, k)]
        assert np.abs(logits.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (tag, pct)


@pytest.mark.parametrize("tag,pct", [("gd_first", 2.4), ("gd_first", 0.7), ("gd_plain", 1.5), ("gd_plain", 2.0),
                                     ("gd_second", 2.4), ("gd_second", 0.7)])
def test_wgan_gp_critic_loss_and_gradients_match_double_backward(tag, pct):
    c, cfg, store, x, y, d = _setup(tag)
    dev = d.cx.device
    g = (y * 0.6 + 0.05 * torch.sin(torch.arange(y.numel(), dtype=torch.float64)).view_as(y))
    lf = torch.tensor([[0.3], [0.8]], dtype=torch.float64)
    # oracle: autograd with create_graph (tf.gradients inside the loss, GAN/multipassGAN-8x.py:1120-1138)
    ctx = ot.TrainContext(store, torch.float64)
    disc, _ = o8.growing_disc(y, x, pct, ctx, cfg)
    gen, _ = o8.growing_disc(g, x, pct, ctx, cfg)
    # refinement networks (gd_second): one gradient norm per (sample, image column), see oracle.wgan_gp_losses
    L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc(t, x, pct, ctx, cfg)[0], y, g, lf,
                          image_side=None if cfg.upsampling_mode == 2 else cfg.tileSizeHigh)
    names = [n for n, t in ctx.leaves.items() if t.requires_grad]
    grads = torch.autograd.grad(L["disc_loss"], [ctx.leaves[n] for n in names], allow_unused=True)
    want = {n: (gr.numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape))) for n, gr in zip(names, grads)}
    # GPU
    out = d.critic_step(x.float().to(dev), y.float().to(dev), g.float().to(dev), pct, lf)
    got = d.grads()
    host = out.cpu().numpy()
    assert abs(host[0] - float(L["disc_loss"])) < 2e-4 * max(1.0, abs(float(L["disc_loss"]))), (host, float(L["disc_loss"]))
    assert abs(host[1] - float(L["grad_penalty"])) < 2e-4 * max(1.0, float(L["grad_penalty"]))
    worst = 0.0
    for n in names:
        if np.abs(want[n]).max() == 0.0:
            assert np.abs(got[n]).max() < 1e-6, n  # d_cfromDensity1 in firstNNArch mode: never reaches the logits
            continue
        worst = max(worst, _rel(got[n], want[n]))
        assert _rel(got[n], want[n]) < 2e-3, (n, _rel(got[n], want[n]))
    assert worst > 0.0


@pytest.mark.parametrize("tag", ["gd_first", "gd_second"])
def test_critic_input_gradient_for_the_generator_step(tag):
    """g_loss_d = mean(-D(G(x))) (:1117): the gradient the generator receives through the critic."""
    c, cfg, store, x, y, d = _setup(tag)
    dev = d.cx.device
    pct = 1.6
    yy = y.clone().requires_grad_(True)
    logits, _ = o8.growing_disc(yy, x, pct, og.Context(store, torch.float64), cfg)
    (-logits).mean().backward()
    d.cx.st = torch.cuda.current_stream(dev).cuda_stream
    d.refresh()
    lg, sv = d.forward(x.float().to(dev), y.float().to(dev), pct)
    dl = torch.full_like(lg, -1.0 / lg.shape[0])
    dxin = d.backward(sv, dl, need_input_grad=True, param_grads=False)
    got = dxin[..., 1].reshape(y.shape).cpu().numpy()
    assert _rel(got, yy.grad.numpy()) < 1e-4


def test_staged_adam_and_weight_ema():
    c, cfg, store, x, y, d = _setup("gd_first")
    dev = d.cx.device
    g = y * 0.5
    lf = torch.tensor([[0.5], [0.25]], dtype=torch.float64)
    names = sorted(store.values)
    opt = t8.StagedAdam(d.cx, d.ps, [1e-3, 2e-3, 3e-3], beta1=0.0, beta2=0.99)
    ema = t8.WeightEMA(d.ps, 0.999)
    ref_vals = {k: np.array(v, np.float64) for k, v in store.values.items()}
    ref_opts = [ot.Adam(lr, 0.0, 0.99) for lr in (1e-3, 2e-3, 3e-3)]
    shadow = o8.ema_init(ref_vals)
    for z, pct in ((0, 0.5), (0, 0.9), (1, 1.5), (2, 2.5)):
        d.critic_step(x.float().to(dev), y.float().to(dev), g.float().to(dev), pct, lf)
        opt.step(z)
        ema.update(opt.state[z]["mask"])
        # oracle step on the same loss with the stage's variable list
        st = og.VarStore(seed=c["seed"])
        st.values = {k: v.astype(np.float32) for k, v in ref_vals.items()}
        ctx = ot.TrainContext(st, torch.float64)
        for k in ctx.leaves:
            pass
        disc, _ = o8.growing_disc(y, x, pct, ctx, cfg)
        gen, _ = o8.growing_disc(g, x, pct, ctx, cfg)
        L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc(t, x, pct, ctx, cfg)[0], y, g, lf)
        sel = o8.stage_variables(names, z)
        grads = torch.autograd.grad(L["disc_loss"], [ctx.leaves[n] for n in sel], allow_unused=True)
        gd = {n: (gr.numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape))) for n, gr in zip(sel, grads)}
        vals32 = {n: ref_vals[n].astype(np.float32) for n in sel}
        ref_opts[z].step(vals32, gd)
        for n in sel:
            ref_vals[n] = vals32[n].astype(np.float64)
        o8.ema_update(shadow, {n: ref_vals[n] for n in sel}, 0.999)
    got = d.ps.export()
    got_ema = ema.export()
    for n in names:
        assert np.abs(got[n] - ref_vals[n]).max() < 2e-4, (n, float(np.abs(got[n] - ref_vals[n]).max()))
        assert np.abs(got_ema[n] - shadow[n]).max() < 1e-5, n
    # stage 0 / 1 never touched the first blocks: those variables are still at their initial values after the z<2 steps
    # (checked implicitly above: the oracle only updated `sel`)


def _gen_setup(tag="gg_first"):
    c = json.loads(str(GOLD[tag + "_cfg"]))
    mode, fs = c.get("upsampling_mode", 2), c.get("filterSize", 3)
    cfg = o8.Cfg8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, c.get("first_nn_arch", True), upsampling_mode=mode)
    store = og.VarStore(seed=c["seed"])
    x = torch.from_numpy(GOLD[tag + "_x"]).double()
    o8.growing_gen_train(x, 1.0, og.Context(store, torch.float64), cfg)
    g = t8.GrowingGen(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, batch=2, values=store.values, upsampling_mode=mode)
    assert {n for n, *_ in g.ps.specs} == set(store.values)
    return c, cfg, store, x, g


@pytest.mark.parametrize("tag", ["gg_first", "gg_second"])
def test_growing_gen_training_forward_matches_the_reference_vectors(tag):
    c, cfg, store, x, g = _gen_setup(tag)
    dev = g.cx.device
    g.cx.st = torch.cuda.current_stream(dev).cuda_stream
    g.refresh()
    for k, pct in enumerate(c["percentages"]):
        out, _ = g.forward(x.float().to(dev), pct)
        ref = GOLD["%s_p%d_out" % (tag, k)]
        assert np.abs(out.cpu().numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), pct


def _gen_autograd(store, x, pct, cfg, dtype):
    ctx = ot.TrainContext(store, dtype)
    out = o8.growing_gen_train(x.to(dtype), pct, ctx, cfg)
    wgt = torch.cos(torch.arange(out.numel(), dtype=torch.float64) * 0.37).view_as(out)
    names = [n for n, t in ctx.leaves.items() if t.requires_grad]
    grads = torch.autograd.grad((out * wgt.to(dtype)).sum(), [ctx.leaves[n] for n in names], allow_unused=True)
    return names, wgt, {n: (gr.double().numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape)))
                        for n, gr in zip(names, grads)}


@pytest.mark.parametrize("tag,pct", [("gg_first", 0.6), ("gg_first", 1.7), ("gg_first", 2.5), ("gg_second", 0.6), ("gg_second", 2.5)])
def test_growing_gen_backward_matches_autograd(tag, pct):
    """Parameter gradients of the training-mode generator (30 convs, 20 pixel norms, ReLU kinks) against fp64 autograd. The
    earliest layers see the rounding of everything above them (a ReLU input near zero may change sign between fp32 and fp64),
    so each variable is held to 3e-3 or to 4x the distance of torch's OWN fp32 autograd from the fp64 result."""
    c, cfg, store, x, g = _gen_setup(tag)
    dev = g.cx.device
    names, wgt, want = _gen_autograd(store, x, pct, cfg, torch.float64)
    _, _, want32 = _gen_autograd(store, x, pct, cfg, torch.float32)
    g.cx.st = torch.cuda.current_stream(dev).cuda_stream
    g.refresh()
    g.ps.gw.zero_()
    gen, sv = g.forward(x.float().to(dev), pct)
    g.backward(sv, wgt.float().to(dev).contiguous())
    g.cx.call("mul", g.ps.g, g.ps.gw, g.ps.scale, g.ps.total, g.cx.st)
    got = g.grads()
    tight = easy = live = 0
    for n in names:
        if np.abs(want[n]).max() < 1e-12:  # stages that are not blended in yet receive no gradient
            assert np.abs(got[n]).max() < 1e-5, n
            continue
        live += 1
        r32 = _rel(want32[n], want[n])
        tol = max(3e-3, 4.0 * r32)
        assert _rel(got[n], want[n]) < tol, (n, _rel(got[n], want[n]), tol)
        easy += r32 < 7.5e-4             # variables torch's fp32 autograd itself resolves well ...
        tight += r32 < 7.5e-4 and _rel(got[n], want[n]) < 3e-3   # ... must be within 3e-3 here
    assert live > 0 and tight == easy, (tight, easy, live)


@pytest.mark.parametrize("tag", ["gg_first", "gg_second"])
def test_trainer8x_critic_and_generator_steps_track_the_oracle(tag):
    """Two loop bodies (critic step then generator step, GAN/multipassGAN-8x.py:1898-2075 without the temporal terms) at two
    growing stages: losses and every updated variable against the fp64 oracle with the same staged Adam.  gg_first = the first
    network's graph, gg_second = the refinement network's (x_in / y_in wiring, no resampling, row-wise gradient penalty)."""
    c = json.loads(str(GOLD[tag + "_cfg"]))
    mode, fs = c.get("upsampling_mode", 2), c.get("filterSize", 3)
    first = c.get("first_nn_arch", True)
    cfg = o8.Cfg8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, first, upsampling_mode=mode)
    S = cfg.tileSizeHigh
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.random((2, cfg.tileSizeLow ** 2 * cfg.n_inputChannels))).double()
    y = torch.from_numpy(rng.random((2, S * S * (1 if mode == 2 else 2)))).double()
    x_in, y_in = (x, y) if mode == 2 else o8.refine_input(x, y, cfg)
    side = None if mode == 2 else S
    store = og.VarStore(seed=7)
    o8.growing_gen_train(x_in, 1.0, og.Context(store, torch.float64), cfg)
    o8.growing_disc(y_in, x, 1.0, og.Context(store, torch.float64), cfg)
    tr = t8.Trainer8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, batch=2, learning_rate=1e-3, values=store.values,
                      upsampling_mode=mode)
    dev = tr.cx.device
    ref_vals = {k: np.array(v, np.float64) for k, v in store.values.items()}
    g_names = sorted(k for k in ref_vals if k.startswith("generator/"))
    d_names = sorted(k for k in ref_vals if k.startswith("spatial-disc/"))
    assert {n for n, *_ in tr.gen.ps.specs} == set(g_names) and {n for n, *_ in tr.disc.ps.specs} == set(d_names)
    og_opt = [ot.Adam(1e-3, 0.0, 0.99) for _ in range(3)]
    od_opt = [ot.Adam(1e-3, 0.0, 0.99) for _ in range(3)]
    lf = torch.tensor([[0.35], [0.6]], dtype=torch.float64)
    xf, yf = x.float().to(dev), y.float().to(dev)

    def oracle_ctx():
        st = og.VarStore(seed=7)
        st.values = {k: v.astype(np.float32) for k, v in ref_vals.items()}
        return ot.TrainContext(st, torch.float64)

    def apply(opt, sel, loss, ctx):
        grads = torch.autograd.grad(loss, [ctx.leaves[n] for n in sel], allow_unused=True)
        gd = {n: (gr.numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape))) for n, gr in zip(sel, grads)}
        v32 = {n: ref_vals[n].astype(np.float32) for n in sel}
        opt.step(v32, gd)
        for n in sel:
            ref_vals[n] = v32[n].astype(np.float64)

    for z, pct in ((0, 0.7), (1, 1.4)):
        # critic step
        ctx = oracle_ctx()
        gen_y = o8.growing_gen_train(x_in, pct, ctx, cfg).detach()
        disc, _ = o8.growing_disc(y_in, x, pct, ctx, cfg)
        gen, _ = o8.growing_disc(gen_y, x, pct, ctx, cfg)
        L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc(t, x, pct, ctx, cfg)[0], y_in, gen_y, lf, image_side=side)
        got = tr.disc_step(xf, yf, pct, z, lf).cpu().numpy()
        assert abs(got[0] - float(L["disc_loss"].detach())) < 5e-4 * max(1.0, abs(float(L["disc_loss"].detach())))
        apply(od_opt[z], o8.stage_variables(d_names, z), L["disc_loss"], ctx)
        # generator step
        ctx = oracle_ctx()
        gen_y = o8.growing_gen_train(x_in, pct, ctx, cfg)
        gen, _ = o8.growing_disc(gen_y, x, pct, ctx, cfg)
        g_loss = (-gen).mean() + 1.0 * (y_in - gen_y).abs().mean()
        gl = tr.gen_step(xf, yf, pct, z).cpu().numpy()
        assert abs(gl[0] + gl[1] - float(g_loss.detach())) < 5e-4 * max(1.0, abs(float(g_loss.detach())))
        apply(og_opt[z], o8.stage_variables(g_names, z), g_loss, ctx)
    got_g, got_d = tr.gen.ps.export(), tr.disc.ps.export()
    for n in g_names:
        assert np.abs(got_g[n] - ref_vals[n]).max() < 5e-4, (n, float(np.abs(got_g[n] - ref_vals[n]).max()))
    for n in d_names:
        assert np.abs(got_d[n] - ref_vals[n]).max() < 5e-4, (n, float(np.abs(got_d[n] - ref_vals[n]).max()))


def _assert_same_up_to_summation_order(a, b, lr=1e-3):
    """Two runs of the same step differ by the order of the atomics in the filter gradients (1e-7 relative). Adam turns that
    into 1e-10 steps -- except for an element whose gradient is itself rounding noise, where the sign of the step is arbitrary:
    at most a handful of elements may differ by up to one Adam step (2 * lr), everything else must agree."""
    d = np.concatenate([np.abs(a[n] - b[n]).ravel() for n in a])
    assert float(d.max()) <= 2.5 * lr, float(d.max())
    assert int((d >= 2e-6).sum()) <= max(3, d.size // 20000), (int((d >= 2e-6).sum()), d.size)


def _loop_trainer(seed=7, values=None):
    return t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2, learning_rate=1e-3, values=values, seed=seed)


def _batches(dev, L=4, C=6, B=2, seed=5):
    g = torch.Generator(device="cpu").manual_seed(seed)
    seen = []

    def batches(upres):
        seen.append(upres)
        return (torch.rand((B, L * L * C), generator=g).to(dev), torch.rand((B, (L * upres) ** 2), generator=g).to(dev))
    return batches, seen


def test_trainer8x_training_loop_grows_saves_and_resumes(tmp_path):
    """Trainer8x.train over a whole (tiny) schedule: the stage's data resolution reaches `batches`, growing events save the
    model, the per-stage optimizers leave later stages untouched, the learning rate decays, and a run resumed from the saved
    checkpoint continues identically (variables, optimizer moments, moving averages)."""
    from mpgan_b200 import schedule8x as S8, tfckpt
    np.random.seed(3)
    tr = _loop_trainer()
    dev = tr.cx.device
    v0 = tr.values()
    sch = S8.GrowthSchedule(stageIter=2, decayIter=2, upRes=8, upsampling_mode=2)
    batches, seen = _batches(dev)
    grown, logs = [], []
    # stage 1 only (4 iterations, optimizers of index 0): genBlock8 / dBlock8 stay at their initial values
    hist = tr.train(batches, sch, zero_density=False, max_iters=4, log=logs.append, log_interval=1)
    assert seen == [2] * 8 and len(hist) == 4 and all(np.isfinite(h[1:]).all() for h in hist)
    v1 = tr.values()
    moved = {n for n in v0 if not np.array_equal(v0[n], v1[n])}
    assert moved and not any(("genBlock8/" in n or "dBlock8/" in n or "dBlock4/" in n or "genBlock4/g_c" in n) for n in moved), sorted(moved)
    assert any("genBlock2/" in n for n in moved) and any("dBlock2/" in n for n in moved)
    # the whole schedule on a fresh trainer: 14 iterations, two growing events
    np.random.seed(3)
    tr = _loop_trainer()
    batches, seen = _batches(dev)
    d = str(tmp_path / "test_0000")
    hist = tr.train(batches, sch, save_dir=d, saveInterval=5, on_grow=grown.append, zero_density=True, log_interval=1)
    assert grown == [4, 8] and len(hist) == 14 and seen[0] == 2 and seen[-1] == 8 and sorted(set(seen)) == [2, 4, 8]
    assert tr.opt_g.lrs[0] == pytest.approx(S8.polynomial_decay(1e-3, 2, 2, 5e-5, 1.1))
    assert [st["t"] for st in tr.opt_g.state] == [4, 4, 6]
    saved = sorted(f for f in os.listdir(d) if f.endswith(".index"))
    assert saved[:2] == ["model_0000.ckpt.index", "model_0001.ckpt.index"] and len(saved) == 2 * tr.save_no and tr.save_no >= 4
    # the moving-average file carries the generator's shadow values under the variable names multipassGAN-out.py restores
    last = tr.save(d)
    ema = tfckpt.read_checkpoint(os.path.join(d, "model_ema_%04d.ckpt" % last), verify_data=True)
    sh = tr.ema.export()
    assert all(np.array_equal(ema[n], sh[n]) for n in sh) and any(not np.array_equal(sh[n], tr.values()[n]) for n in sh)
    # resume: a fresh trainer restored from the files continues exactly like the original
    tr2 = _loop_trainer(seed=99)
    tr2.load(d, last)
    xs, ys = torch.rand((2, 96), device=dev), torch.rand((2, 1024), device=dev)
    lf = torch.tensor([[0.25], [0.5]])
    for t in (tr, tr2):
        t.opt_g.lrs = t.opt_d.lrs = [1e-3] * 3
        t.disc_step(xs, ys, 2.5, 2, lf)
        t.gen_step(xs, ys, 2.5, 2)
    a, b = tr.values(), tr2.values()   # (filter gradients are summed with atomics: equal up to the summation order)
    _assert_same_up_to_summation_order(a, b)
    assert float((tr.ema.shadow - tr2.ema.shadow).abs().max()) < 1e-7


def test_trainer8x_target_rows_are_the_nearest_resize():
    tr = _loop_trainer()
    dev = tr.cx.device
    tr.cx.st = torch.cuda.current_stream(dev).cuda_stream
    y = torch.rand((2, 64), device=dev)                       # 8x8 tiles (currentUpres 2 at L = 4) -> 32x32
    got = tr.target_rows(y).view(2, 32, 32)
    want = y.view(2, 8, 8).repeat_interleave(4, 1).repeat_interleave(4, 2)
    assert torch.equal(got, want)
    full = torch.rand((2, 1024), device=dev)
    assert tr.target_rows(full) is full
    with pytest.raises(ValueError):
        tr.target_rows(torch.rand((2, 60), device=dev))


def test_trainer8x_refinement_network_training_loop():
    """The second network's shipped schedule (GAN/example_run_training.py:7: upsamplingMode 1, stageIter 1): the data is at 8x
    from the first iteration, every iteration uses the last stage's optimizers, the blend reaches 2 after the first one."""
    from mpgan_b200 import schedule8x as S8
    np.random.seed(11)
    tr = t8.Trainer8x(2, 8, 4, 32, 32, 5, batch=2, learning_rate=1e-3, upsampling_mode=1)
    dev = tr.cx.device
    g = torch.Generator(device="cpu").manual_seed(2)
    seen = []

    def batches(upres):
        seen.append(upres)
        return torch.rand((2, 2 * 2 * 4), generator=g).to(dev), torch.rand((2, 16 * 16 * 2), generator=g).to(dev)

    v0 = tr.values()
    sch = S8.GrowthSchedule(stageIter=1, decayIter=2, upRes=8, upsampling_mode=1)
    hist = tr.train(batches, sch, add_adj_idcs=False, log_interval=1)
    assert len(hist) == 8 and seen == [8] * 16 and all(np.isfinite(h[1:]).all() for h in hist)
    assert [st["t"] for st in tr.opt_g.state] == [0, 0, 8] and [st["t"] for st in tr.opt_d.state] == [0, 0, 8]
    v1 = tr.values()
    for n in ("generator/g_cA_1/weight", "generator/genBlock2/g_cB_first/weight", "generator/genBlock4/g_cdensOut4/weight",
              "spatial-disc/dBlock4/d_cA4/weight", "spatial-disc/d_cB1/weight", "spatial-disc/d_l61/weight"):
        assert not np.array_equal(v0[n], v1[n]), n
    # reference behaviour, reproduced: without growing events the blend counter stops at 2 (schedule8x doc, trace m1_*), so
    # the 8x stage of either network is never blended in and receives no gradient under this schedule
    assert max(h_[0] for h_ in hist) == 7 and all(st.percentage <= 2.0 for st in sch)
    for n in ("generator/genBlock8/g_cdensOut8/weight", "spatial-disc/dBlock8/d_cA8/weight"):
        assert np.array_equal(v0[n], v1[n]), n
    with pytest.raises(ValueError):
        tr.disc_step(torch.rand((2, 16), device=dev), torch.rand((2, 256), device=dev), 2.0, 2, torch.rand(2, 1))


def test_stage_batches_feed_the_loop_from_device_resident_tile_samplers():
    """StageBatches: one TileSampler per data resolution (the reference re-creates its TileCreator at every growing event,
    :1916-1960), tiles gathered on the device, target tiles nearest-resized to the full size inside the loop."""
    import random
    from mpgan_b200 import schedule8x as S8
    from mpgan_b200.tilesampler import TileSampler
    np.random.seed(5)
    tr = _loop_trainer()
    dev = tr.cx.device
    rng = np.random.default_rng(9)
    low = rng.random((6, 1, 8, 8, 6), dtype=np.float32) + 0.1
    samplers = {}
    for u in (2, 4, 8):
        samplers[u] = TileSampler(4, u, densityMinimum=0.02, device=dev, rng=random.Random(u))
        samplers[u].add_data(low, rng.random((6, 1, 8 * u, 8 * u, 1), dtype=np.float32))
    sb = t8.StageBatches(samplers, 2)
    xs, ys = sb(4)
    assert xs.shape == (2, 4 * 4 * 6) and ys.shape == (2, 16 * 16) and xs.device == dev
    hist = tr.train(sb, S8.GrowthSchedule(stageIter=1, decayIter=1), log_interval=1)
    assert len(hist) == 7 and all(np.isfinite(h[1:]).all() for h in hist)
    with pytest.raises(KeyError):
        t8.StageBatches({2: samplers[2]}, 2)(4)


def test_trained_checkpoint_applies_through_the_out_pipeline_graph(tmp_path):
    """train -> model_ema_%04d.ckpt -> the restore rule of multipassGAN-out.py (:367-386, tfckpt.load_generator_weights) -> the
    apply engine: at percentage 3 every stage is fully blended in, so the training-mode generator (per-stage density outputs,
    lerp) and the apply graph (`output=True`, last stage only) are the same function of the same moving-average weights."""
    from mpgan_b200 import engine, graph as G, networks as N, pipeline as P, schedule8x as S8, tfckpt
    np.random.seed(1)
    tr = _loop_trainer()
    dev = tr.cx.device
    batches, _ = _batches(dev)
    d = str(tmp_path / "test_0003")
    tr.train(batches, S8.GrowthSchedule(stageIter=1, decayIter=1), save_dir=d, saveInterval=7, zero_density=False)
    no = tr.save(d)
    spec = P.NetSpec(use_res_net=True, add_adj_idcs=True, startFms=32, maxFms=32, filterSize=3, first_nn_arch=True)
    G.reset_default_graph()
    out = P.build_out_graph(1, spec, N.config_out(4, upRes=8))
    names = list(G.get_default_graph().variables)           # gen_1/generator/... : the names the graph built from the flags needs
    w = tfckpt.load_generator_weights(os.path.join(d, "model_ema_%04d.ckpt" % no), sorted(names), "gen_1")
    sh = tr.ema.export()
    assert all(np.array_equal(w[n], sh[n[len("gen_1/"):]]) for n in w) and len(w) >= 60
    x = torch.rand((2, 4 * 4 * 6), device=dev)
    net = engine.CompiledNet(out, w, 2, precision="fp32")
    applied = net.run({"x": x}).float().cpu().numpy()
    net.close()
    # the trainer's own forward on the moving averages
    keep = tr.gen.ps.v.clone()
    tr.gen.ps.v.copy_(tr.ema.shadow)
    tr.cx.st = torch.cuda.current_stream(dev).cuda_stream
    tr.gen.refresh()
    trained, _ = tr.gen.forward(x, 3.0)
    tr.gen.ps.v.copy_(keep)
    ref = trained.cpu().numpy()
    assert np.abs(applied - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------ temporal critic (growing_disc_tempo, lambda_t)
def _tempo_setup(tag, batch=2):
    c = json.loads(str(GOLD[tag + "_cfg"]))
    mode = c.get("upsampling_mode", 2)
    cfg = o8.Cfg8x(c["L"], c["u"], 4, c["start_fms"], c["max_fms"], c["filterSize"], c["first_nn_arch"], upsampling_mode=mode)
    store = og.VarStore(seed=c["seed"])
    fr = torch.from_numpy(GOLD[tag + "_frames"]).double().reshape(2, -1)
    o8.growing_disc_tempo(fr, 1.0, og.Context(store, torch.float64), cfg)
    d = t8.GrowingDisc(c["L"], c["u"], 4, c["start_fms"], c["max_fms"], c["filterSize"], c["first_nn_arch"], batch=batch,
                       values=store.values, upsampling_mode=mode, kind="tempo")
    assert {n for n, *_ in d.ps.specs} == set(store.values)
    return c, cfg, store, fr, d


@pytest.mark.parametrize("tag", ["gt_first", "gt_second"])
def test_temporal_critic_forward_matches_the_reference_vectors(tag):
    c, cfg, store, fr, d = _tempo_setup(tag)
    dev = d.cx.device
    d.cx.st = torch.cuda.current_stream(dev).cuda_stream
    d.refresh()
    S = cfg.tileSizeHigh
    for k, pct in enumerate(c["percentages"]):
        logits, _ = d.forward_from_input(fr.float().to(dev).view(2, S, S, 3), pct)
        ref = GOLD["%s_p%d_logits" % (tag, k)]
        assert np.abs(logits.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (tag, pct)


@pytest.mark.parametrize("tag,pct", [("gt_first", 2.4), ("gt_first", 0.7), ("gt_second", 2.4), ("gt_second", 1.5)])
def test_temporal_critic_wgan_gp_loss_and_gradients_match_double_backward(tag, pct):
    """t_disc_loss of :1262-1289: one gradient norm per (sample, frame)."""
    c, cfg, store, fr, d = _tempo_setup(tag)
    dev = d.cx.device
    g = fr * 0.6 + 0.05 * torch.sin(torch.arange(fr.numel(), dtype=torch.float64)).view_as(fr)
    lf = torch.tensor([[0.3], [0.8]], dtype=torch.float64)
    ctx = ot.TrainContext(store, torch.float64)
    disc = o8.growing_disc_tempo(fr, pct, ctx, cfg)
    gen = o8.growing_disc_tempo(g, pct, ctx, cfg)
    L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc_tempo(t, pct, ctx, cfg), fr, g, lf, frames=3)
    names = [n for n, t in ctx.leaves.items() if t.requires_grad]
    grads = torch.autograd.grad(L["disc_loss"], [ctx.leaves[n] for n in names], allow_unused=True)
    want = {n: (gr.numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape))) for n, gr in zip(names, grads)}
    out = d.critic_step_frames(fr.float().to(dev), g.float().to(dev), pct, lf).cpu().numpy()
    got = d.grads()
    assert abs(out[0] - float(L["disc_loss"].detach())) < 2e-4 * max(1.0, abs(float(L["disc_loss"].detach()))), (out, L["disc_loss"])
    assert abs(out[1] - float(L["grad_penalty"].detach())) < 2e-4 * max(1.0, float(L["grad_penalty"].detach()))
    worst = 0.0
    for n in names:
        if np.abs(want[n]).max() == 0.0:
            assert np.abs(got[n]).max() < 1e-6, n
            continue
        worst = max(worst, _rel(got[n], want[n]))
        assert _rel(got[n], want[n]) < 2e-3, (n, _rel(got[n], want[n]))
    assert worst > 0.0


def test_tensor_resample_kernels_match_the_reference_vectors_and_autograd():
    """mpg_train_resample_fwd against the reference's own tensorResample (vectors incl. out-of-range positions), _bwd against
    autograd of the torch restatement."""
    from mpgan_b200 import capi
    val, pos = GOLD["resample_value"], GOLD["resample_pos"]
    n, hh, ww, c = val.shape
    h = capi.default_handle(0)
    v, p = torch.from_numpy(val).cuda(), torch.from_numpy(pos).cuda()
    out = torch.empty_like(v)
    capi.train_call("resample_fwd", h, v, p, out, n, hh, ww, c, 0)
    assert np.abs(out.cpu().numpy() - GOLD["resample_out"]).max() < 1e-5
    vv = torch.from_numpy(val).double().requires_grad_(True)
    wgt = torch.cos(torch.arange(val.size, dtype=torch.float64) * 0.61).view(val.shape)
    (o8.tensor_resample(vv, torch.from_numpy(pos).double()) * wgt).sum().backward()
    dv = torch.zeros_like(v)
    capi.train_call("resample_bwd", h, wgt.float().cuda().contiguous(), p, dv, n, hh, ww, c, 0)
    assert np.abs(dv.cpu().numpy() - vv.grad.numpy()).max() < 1e-5


@pytest.mark.parametrize("tag,advected", [("gg_first", False), ("gg_second", False), ("gg_first", True)])
def test_trainer8x_temporal_steps_track_the_oracle(tag, advected):
    """lambda_t 1.0 as in both shipped commands (aligned triplets, adv_flag 0): one temporal-critic step (:2001-2013) and one
    generator step whose loss carries kkt * mean(-T(G(x_t))) next to the spatial terms (:2015-2043), against the fp64 oracle
    with the same staged Adam: losses and every variable of the three networks."""
    c = json.loads(str(GOLD[tag + "_cfg"]))
    mode, fs = c.get("upsampling_mode", 2), c.get("filterSize", 3)
    cfg = o8.Cfg8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, c.get("first_nn_arch", True), upsampling_mode=mode)
    S, L_, C = cfg.tileSizeHigh, cfg.tileSizeLow, cfg.n_inputChannels
    ych = 1 if mode == 2 else 2
    # (seed: with default_rng(4) one pre-activation of an otherwise all-zero pixel sits within rounding of 0 -- the kernels' ReLU
    #  mask then differs from the fp64 one in ONE element, and the pixel norm of a near-zero vector multiplies that element's
    #  gradient by 1e4: 8 % on two bias gradients from a measure-zero event. Not a property of the step, so not tested here.)
    rng = np.random.default_rng(6)
    x = torch.from_numpy(rng.random((2, L_ * L_ * C))).double()
    y = torch.from_numpy(rng.random((2, S * S * ych))).double()
    xt = torch.from_numpy(rng.random((6, L_ * L_ * C))).double()          # 2 samples x 3 frames
    yt = torch.from_numpy(rng.random((6, S * S * ych))).double()
    conv = (lambda a, b: (a, b)) if mode == 2 else (lambda a, b: o8.refine_input(a, b, cfg))
    x_in, y_in = conv(x, y)
    xt_in, yt_in = conv(xt, yt)
    side = None if mode == 2 else S

    # advected: adv_flag 1 / adv_mode 0 of the shipped commands -- both the generated and the target frames are re-sampled at
    # given positions (tensorResample, :1195-1197, 1241-1242) before they become the critic's three channels
    pos = None
    if advected:
        base = np.stack(np.meshgrid(np.arange(S) + 0.5, np.arange(S) + 0.5, indexing="ij"), axis=-1)[None]
        pos = torch.from_numpy((base + rng.normal(0.0, 0.8, (6, S, S, 2))).reshape(6, -1)).double()

    def frames(rows):                                                    # :1213-1214
        if pos is not None:
            rows = o8.tensor_resample(rows.reshape(-1, S, S, 1), pos.to(rows.dtype).reshape(-1, S, S, 2)).reshape(-1, S * S)
        return rows.reshape(-1, 3, S * S).permute(0, 2, 1).reshape(-1, S * S * 3)

    store = og.VarStore(seed=7)
    o8.growing_gen_train(x_in, 1.0, og.Context(store, torch.float64), cfg)
    o8.growing_disc(y_in, x, 1.0, og.Context(store, torch.float64), cfg)
    o8.growing_disc_tempo(frames(yt_in), 1.0, og.Context(store, torch.float64), cfg)
    tr = t8.Trainer8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], fs, batch=2, learning_rate=1e-3, values=store.values,
                      upsampling_mode=mode, lambda_t=1.0)
    dev = tr.cx.device
    ref_vals = {k: np.array(v, np.float64) for k, v in store.values.items()}
    g_names = sorted(k for k in ref_vals if k.startswith("generator/"))
    t_names = sorted(k for k in ref_vals if k.startswith("tempo-disc/"))
    assert {n for n, *_ in tr.tdisc.ps.specs} == set(t_names) and len(t_names) >= 20
    z, pct = 1, 1.4
    lf = torch.tensor([[0.35], [0.6]], dtype=torch.float64)

    live, want_g = {}, {}

    def oracle_ctx():
        st = og.VarStore(seed=7)
        st.values = {k: v.astype(np.float32) for k, v in ref_vals.items()}
        return ot.TrainContext(st, torch.float64)

    def apply(sel, loss, ctx):
        grads = torch.autograd.grad(loss, [ctx.leaves[n] for n in sel], allow_unused=True)
        gd = {n: (gr.numpy() if gr is not None else np.zeros(tuple(ctx.leaves[n].shape))) for n, gr in zip(sel, grads)}
        v32 = {n: ref_vals[n].astype(np.float32) for n in sel}
        ot.Adam(1e-3, 0.0, 0.99).step(v32, gd)
        for n in sel:
            ref_vals[n] = v32[n].astype(np.float64)
            # the first Adam step moves every element by lr * sign(gradient): elements whose gradient is at the rounding level of
            # the fp32 kernels (relative to the variable's largest) may legitimately step the other way
            live[n] = np.abs(gd[n]) > 2e-2 * max(np.abs(gd[n]).max(), 1e-30)
            want_g[n] = gd[n]

    # temporal critic step
    ctx = oracle_ctx()
    gen_ts = o8.growing_gen_train(xt_in, pct, ctx, cfg).detach()
    real, fake = frames(yt_in), frames(gen_ts)
    disc_s = o8.growing_disc_tempo(real, pct, ctx, cfg)
    gen_s = o8.growing_disc_tempo(fake, pct, ctx, cfg)
    Lt = o8.wgan_gp_losses(disc_s, gen_s, lambda t: o8.growing_disc_tempo(t, pct, ctx, cfg), real, fake, lf, frames=3)
    pos_d = pos.float().to(dev).contiguous() if pos is not None else None
    got = tr.t_disc_step(xt.float().to(dev), yt.float().to(dev), pct, z, lf, pos_d).cpu().numpy()
    assert abs(got[0] - float(Lt["disc_loss"].detach())) < 5e-4 * max(1.0, abs(float(Lt["disc_loss"].detach())))
    apply(o8.stage_variables(t_names, z), Lt["disc_loss"], ctx)
    # generator step: g_loss_d + l1 + kkt * g_loss_t
    vals_before_gen = {k: v.copy() for k, v in ref_vals.items()}
    ctx = oracle_ctx()
    gen_y = o8.growing_gen_train(x_in, pct, ctx, cfg)
    gen, _ = o8.growing_disc(gen_y, x, pct, ctx, cfg)
    gen_ts = o8.growing_gen_train(xt_in, pct, ctx, cfg)
    g_loss_t = (-o8.growing_disc_tempo(frames(gen_ts), pct, ctx, cfg)).mean()
    g_loss = (-gen).mean() + 1.0 * (y_in - gen_y).abs().mean() + 1.0 * g_loss_t
    gl = tr.gen_step(x.float().to(dev), y.float().to(dev), pct, z, xt.float().to(dev), yt.float().to(dev), pos_d).cpu().numpy()
    assert abs(gl[2] - float(g_loss_t.detach())) < 5e-4 * max(1.0, abs(float(g_loss_t.detach())))
    assert abs(gl[0] + gl[1] + gl[2] - float(g_loss.detach())) < 5e-4 * max(1.0, abs(float(g_loss.detach())))
    apply(o8.stage_variables(g_names, z), g_loss, ctx)
    got_g, got_t = tr.gen.ps.export(), tr.tdisc.ps.export()
    # the gradients the two optimizers consumed (still in ps.g), variable by variable. The generator's earliest layers see the
    # rounding of everything above them twice now (two passes through ~30 convs / 20 pixel norms with ReLU kinks): as in
    # test_growing_gen_backward_matches_autograd each variable is held to 5e-3 or to 4x the distance of torch's OWN fp32
    # autograd (same losses, same weights) from the fp64 result
    for n, gq in tr.tdisc.grads().items():
        if n in want_g and np.abs(want_g[n]).max() > 0:
            assert _rel(gq, want_g[n]) < 5e-3, (n, _rel(gq, want_g[n]))
    st32 = og.VarStore(seed=7)
    st32.values = {k: v.astype(np.float32) for k, v in vals_before_gen.items()}
    c32 = ot.TrainContext(st32, torch.float32)
    f32 = lambda t: t.float()
    gy32 = o8.growing_gen_train(f32(x_in), pct, c32, cfg)
    gs32, _ = o8.growing_disc(gy32, f32(x), pct, c32, cfg)
    gt32 = o8.growing_gen_train(f32(xt_in), pct, c32, cfg)
    loss32 = (-gs32).mean() + (f32(y_in) - gy32).abs().mean() + (-o8.growing_disc_tempo(frames(gt32), pct, c32, cfg)).mean()
    sel = o8.stage_variables(g_names, z)
    g32 = dict(zip(sel, torch.autograd.grad(loss32, [c32.leaves[n] for n in sel], allow_unused=True)))
    worst = []
    for n, gq in tr.gen.grads().items():
        if n in want_g and np.abs(want_g[n]).max() > 0:
            r32 = _rel(g32[n].double().numpy(), want_g[n]) if g32[n] is not None else 0.0
            worst.append((_rel(gq, want_g[n]), r32, n))
    print("generator gradients, worst (rel-L2 here, rel-L2 of torch fp32, name):", sorted(worst, reverse=True)[:8])
    for e, r32, n in worst:
        assert e < max(5e-3, 4.0 * r32), (n, e, r32)
    checked = total = 0
    for names_, got_ in ((g_names, got_g), (t_names, got_t)):
        for n in names_:
            m = live.get(n, np.ones(ref_vals[n].shape, bool))
            diff = np.abs(got_[n] - ref_vals[n])
            assert diff[m].max(initial=0.0) < 5e-4, (n, float(diff[m].max()))
            assert diff.max() < 2.1e-3, n                      # a flipped sign is 2 * lr, never more
            checked, total = checked + int(m.sum()), total + m.size
    assert checked > 0.5 * total, (checked, total)


def test_training_loop_with_the_temporal_critic_and_checkpoints(tmp_path):
    """Trainer8x.train with lambda_t > 0 and aligned frame triplets: the temporal critic trains with its own staged
    optimizers, its variables and moments travel in the checkpoint, a restored trainer continues like the original."""
    from mpgan_b200 import schedule8x as S8
    np.random.seed(2)
    tr = t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2, learning_rate=1e-3, seed=7, lambda_t=1.0)
    dev = tr.cx.device
    batches, seen = _batches(dev)
    g = torch.Generator(device="cpu").manual_seed(8)
    tseen = []

    def tempo_batches(upres):
        tseen.append(upres)
        return torch.rand((6, 4 * 4 * 6), generator=g).to(dev), torch.rand((6, (4 * upres) ** 2), generator=g).to(dev)

    t0 = tr.tdisc.ps.export()
    d = str(tmp_path / "test_0007")
    hist = tr.train(batches, S8.GrowthSchedule(stageIter=1, decayIter=1), tempo_batches=tempo_batches, save_dir=d, saveInterval=100,
                    log_interval=1)
    assert len(hist) == 7 and all(np.isfinite(h[1:]).all() for h in hist)
    assert tseen == [2, 2, 2, 2, 4, 4, 4, 4, 8, 8, 8, 8, 8, 8] and [st["t"] for st in tr.opt_t.state] == [2, 2, 3]
    t1 = tr.tdisc.ps.export()
    assert any(not np.array_equal(t0[n], t1[n]) for n in t0)
    no = tr.save(d)
    tr2 = t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2, learning_rate=1e-3, seed=99, lambda_t=1.0)
    tr2.load(d, no)
    xs, ys = torch.rand((2, 96), device=dev), torch.rand((2, 1024), device=dev)
    xt, yt = torch.rand((6, 96), device=dev), torch.rand((6, 1024), device=dev)
    lf = torch.tensor([[0.25], [0.5]])
    for t in (tr, tr2):
        t.opt_g.lrs = t.opt_d.lrs = t.opt_t.lrs = [1e-3] * 3
        t.t_disc_step(xt, yt, 2.5, 2, lf)
        t.gen_step(xs, ys, 2.5, 2, xt, yt)
    a, b = tr.values(), tr2.values()
    assert any(n.startswith("tempo-disc/") for n in a)
    _assert_same_up_to_summation_order(a, b)
    with pytest.raises(ValueError):
        t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2).t_disc_step(xt, yt, 2.5, 2, lf)


# ------------------------------------------------------------------ getTempoinput: three-frame tiles + advected positions
TEMPO = np.load(os.path.join(os.path.dirname(__file__), "golden", "tempotiles.npz"))


def test_semi_lagrangian_positions_kernel_matches_the_reference_vectors():
    """mpg_train_semilagr_pos against getSemiLagrPosBatch / selectRandomTempoTiles executed from the reference
    (tools_wscale/tilecreator_t.py:1341-1413): interpolating (x4, x8) and same-size grids, dt = 0.5 * (+1, 0, -1)."""
    from mpgan_b200 import capi
    h = capi.default_handle(0)
    T, u, C = (int(v) for v in TEMPO["cfg"][:3])
    S = T * u
    xt = torch.from_numpy(TEMPO["xt"]).cuda()
    pos = torch.empty((6, S * S * 2), device="cuda")
    capi.train_call("semilagr_pos", h, xt, pos, 6, T, S, C, 1, 0.5, 3, 0)
    assert np.abs(pos.cpu().numpy() - TEMPO["pos"]).max() < 2e-5
    vel = torch.from_numpy(np.ascontiguousarray(TEMPO["vel"][:3, 0])).cuda()      # rows 0..2 carry dt = 0.5, 0, -0.5
    for side, key in ((24, "pos_up8"), (3, "pos_same")):
        p = torch.empty((3, side, side, 2), device="cuda")
        capi.train_call("semilagr_pos", h, vel, p, 3, 3, side, 3, 0, 0.5, 3, 0)
        assert np.abs(p.cpu().numpy() - TEMPO[key][:3]).max() < 2e-5, key


def _tempo_sampler(dev, u, seed, rng_data):
    import random
    from mpgan_b200.tilesampler import TileSampler
    s = TileSampler(4, u, densityMinimum=0.02, device=dev, rng=random.Random(seed))
    low = rng_data.random((5, 1, 8, 8, 18), dtype=np.float32) + 0.05            # 3 frames x (d, vx, vy, vz, d-, d+)
    low[..., 1:4] -= 0.5
    high = rng_data.random((5, 1, 8 * u, 8 * u, 3), dtype=np.float32)
    s.add_data(low, high)
    return s


def test_tempo_batches_reproduce_select_random_tempo_tiles():
    dev = torch.device("cuda", 0)
    a = _tempo_sampler(dev, 4, 21, np.random.default_rng(50))
    b = _tempo_sampler(dev, 4, 21, np.random.default_rng(50))
    x, y, pos = t8.TempoBatches({4: a}, 6)(4)
    low, high = b.select_random_tiles(2, True, False)
    wx, wy, wpos = o8.tempo_tiles(low.cpu().numpy(), high.cpu().numpy(), 3, 0.5)
    assert x.shape == (6, 4 * 4 * 6) and y.shape == (6, 16 * 16) and pos.shape == (6, 16 * 16 * 2)
    assert np.array_equal(x.cpu().numpy(), wx) and np.array_equal(y.cpu().numpy(), wy)
    assert np.abs(pos.cpu().numpy() - wpos).max() < 2e-5
    with pytest.raises(KeyError):
        t8.TempoBatches({4: a}, 6)(8)


def test_frame_alignment_on_a_coarser_grid_and_its_adjoint():
    """_align for positions on the stage's grid (cur = S / 2): strided pick -> tensorResample -> nearest resize (:1192-1204),
    against the oracle, and _align_bwd is its adjoint (<align(a), b> = <a, align_bwd(b)>)."""
    tr = _loop_trainer()
    dev, S = tr.cx.device, tr.gen.S
    tr.cx.st = torch.cuda.current_stream(dev).cuda_stream
    cur = S // 2
    rng = np.random.default_rng(60)
    a = torch.from_numpy(rng.random((6, S * S), dtype=np.float32)).to(dev)
    base = np.stack(np.meshgrid(np.arange(cur) + 0.5, np.arange(cur) + 0.5, indexing="ij"), axis=-1)[None]
    pos = torch.from_numpy((base + rng.normal(0, 0.9, (6, cur, cur, 2))).astype(np.float32).reshape(6, -1)).to(dev)
    got = tr._align(a, pos)
    picked = a.view(6, S, S)[:, ::2, ::2].double().cpu().unsqueeze(-1)
    want = o8.tensor_resample(picked, pos.double().cpu().view(6, cur, cur, 2))[..., 0]
    want = want.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(6, -1)
    assert float((got.double().cpu() - want).abs().max()) < 1e-5
    bvec = torch.from_numpy(rng.standard_normal((6, S * S)).astype(np.float32)).to(dev)
    lhs = float((got.double() * bvec.double()).sum())
    rhs = float((a.double() * tr._align_bwd(bvec, pos).double()).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
    assert tr._align(a, None) is a


def test_training_loop_on_three_frame_tiles_with_advected_positions():
    """The temporal path end to end: TempoBatches (three-frame tiles + semi-Lagrangian positions at the stage's resolution) ->
    tensorResample alignment -> temporal critic and generator steps, over a growing schedule."""
    from mpgan_b200 import schedule8x as S8
    np.random.seed(4)
    tr = t8.Trainer8x(4, 8, 6, 32, 32, 3, batch=2, learning_rate=1e-3, seed=7, lambda_t=1.0)
    dev = tr.cx.device
    rng = np.random.default_rng(70)
    tb = t8.TempoBatches({u: _tempo_sampler(dev, u, 30 + u, rng) for u in (2, 4, 8)}, 6)
    batches, _ = _batches(dev)
    g0 = tr.gen.ps.export()
    hist = tr.train(batches, S8.GrowthSchedule(stageIter=1, decayIter=1), tempo_batches=tb, log_interval=1)
    assert len(hist) == 7 and all(np.isfinite(h[1:]).all() for h in hist)
    g1 = tr.gen.ps.export()
    assert any(not np.array_equal(g0[n], g1[n]) for n in g0) and [st["t"] for st in tr.opt_t.state] == [2, 2, 3]
