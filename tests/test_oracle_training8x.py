"""8x progressive-growing trainer pieces: the fp64 oracle against golden vectors produced by executing the
reference's own growing_disc / growBlockDisc / lerp (GAN/multipassGAN-8x.py:596-597, 752-866) on the numpy TF1 shim, plus
the staged-variable rule, the EMA and the WGAN-GP loss restatement (CPU only)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import gan as og
from oracle import training as ot
from oracle import training8x as o8

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "growdisc.npz"))


def _cfg(tag):
    c = json.loads(str(GOLD[tag + "_cfg"]))
    return c, o8.Cfg8x(c["L"], c["u"], c["C"], c["start_fms"], c["max_fms"], c["filterSize"], c["first_nn_arch"],
                       upsampling_mode=c.get("upsampling_mode", 2))


def test_growing_disc_oracle_reproduces_the_reference_code():
    for tag in ("gd_first", "gd_plain", "gd_second"):   # gd_second: upsampling_mode 1 (no pooling anywhere)
        c, cfg = _cfg(tag)
        store = og.VarStore(seed=c["seed"])
        x = torch.from_numpy(GOLD[tag + "_x"]).double()
        y = torch.from_numpy(GOLD[tag + "_y"]).double()
        for k, pct in enumerate(c["percentages"]):
            ctx = og.Context(store, torch.float64)
            logits, feats = o8.growing_disc(y, x, pct, ctx, cfg)
            ref = GOLD["%s_p%d_logits" % (tag, k)]
            assert np.abs(logits.numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), (tag, pct)
            for i, f in enumerate(feats):
                key = "%s_p%d_feat%d" % (tag, k, i)
                if key in GOLD.files:
                    assert f.shape == GOLD[key].shape and np.abs(f.numpy() - GOLD[key]).max() < 1e-5, (key,)
                else:
                    s = GOLD[key + "_sums"]
                    assert abs(float(f.sum()) - s[0]) < 1e-3 * max(1.0, abs(s[1])) and abs(float(f.abs().sum()) - s[1]) < 1e-4 * max(1.0, s[1]), key
        # the variables the reference's graph asked for (names AND shapes) are the ones the oracle created
        want = {k: tuple(v) for k, v in json.loads(str(GOLD[tag + "_vars"]))}
        assert {k: tuple(v.shape) for k, v in store.values.items()} == want


def test_stage_variables_follow_the_substring_rule():
    _, cfg = _cfg("gd_first")
    names = [k for k, _ in json.loads(str(GOLD["gd_first_vars"]))]
    z0 = o8.stage_variables(names, 0)
    assert all(("1" in n) or ("2" in n) for n in z0) and any("d_cfromDensity1" in n for n in z0)
    assert not any(n.startswith("spatial-disc/dBlock8") for n in z0)
    assert any("d_l61" in n for n in z0)  # "1" is a substring of d_l61: the FC head trains from the first stage on
    z1 = o8.stage_variables(names, 1)
    assert set(z0) < set(z1) and any("dBlock4" in n for n in z1) and not any("dBlock8/" in n for n in z1)
    assert o8.stage_variables(names, 2) == names


def test_ema_matches_exponential_moving_average():
    v = {"a": np.array([1.0, 2.0]), "b": np.array([[3.0]])}
    sh = o8.ema_init(v)
    v2 = {"a": np.array([2.0, 0.0]), "b": np.array([[5.0]])}
    o8.ema_update(sh, v2, 0.999)
    assert np.allclose(sh["a"], 0.999 * v["a"] + 0.001 * v2["a"]) and np.allclose(sh["b"], 3.0 * 0.999 + 5.0 * 0.001)


def test_wgan_gp_gradient_penalty_is_differentiable_wrt_the_critic():
    """Double backward through growing_disc (tf.gradients inside the loss, :1120-1133): autograd gives finite, non-zero
    parameter gradients, and the penalty of a critic scaled by c has gradient norms scaled by c."""
    c, cfg = _cfg("gd_first")
    store = og.VarStore(seed=c["seed"])
    ctx = ot.TrainContext(store, torch.float64)
    x = torch.from_numpy(GOLD["gd_first_x"]).double()
    y = torch.from_numpy(GOLD["gd_first_y"]).double()
    g = (y * 0.5 + 0.1)
    pct = 2.4
    disc, _ = o8.growing_disc(y, x, pct, ctx, cfg)
    gen, _ = o8.growing_disc(g, x, pct, ctx, cfg)
    lf = torch.tensor([[0.3], [0.8]], dtype=torch.float64)
    L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc(t, x, pct, ctx, cfg)[0], y, g, lf)
    names = [n for n, t in ctx.leaves.items() if t.requires_grad]
    grads = torch.autograd.grad(L["disc_loss"], [ctx.leaves[n] for n in names], allow_unused=True)
    got = {n: gr for n, gr in zip(names, grads) if gr is not None}
    assert any(float(gr.abs().max()) > 0 for gr in got.values()) and all(torch.isfinite(gr).all() for gr in got.values())
    # d_cfromDensity1 feeds only the feature list in firstNNArch mode (cursor quirk): no gradient from the critic losses
    k = "spatial-disc/d_cfromDensity1/weight"
    assert k not in got or float(got[k].abs().max()) == 0.0
    assert L["grad_norms"].shape == (2,) and float(L["grad_penalty"]) > 0


@pytest.mark.parametrize("tag", ["gg_first", "gg_second"])
def test_growing_gen_training_graph_reproduces_the_reference_code(tag):
    """growing_gen with output=False (per-stage density outputs blended with lerp, GAN/multipassGAN-8x.py:700-750): the first
    network (firstNNArch, upsampling_mode 2) and the refinement network (upsampling_mode 1, two head resBlocks, no resampling)."""
    c, cfg = _cfg(tag)
    store = og.VarStore(seed=c["seed"])
    x = torch.from_numpy(GOLD[tag + "_x"]).double()
    for k, pct in enumerate(c["percentages"]):
        out = o8.growing_gen_train(x, pct, og.Context(store, torch.float64), cfg)
        ref = GOLD["%s_p%d_out" % (tag, k)]
        assert out.shape == ref.shape and np.abs(out.numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), pct
    want = {k: tuple(v) for k, v in json.loads(str(GOLD[tag + "_vars"]))}
    assert {k: tuple(v.shape) for k, v in store.values.items()} == want


def test_refinement_gradient_penalty_reduces_over_image_rows():
    """upsampling_mode 1 / 3: the reference's reduce_sum(axis=1) acts on [B, S, S, 1] gradients (:1130), so there is one norm
    per (sample, column); refine_input builds x_in / y_in of :1042-1044, 1061-1062."""
    c, cfg = _cfg("gd_second")
    S = cfg.tileSizeHigh
    store = og.VarStore(seed=c["seed"])
    ctx = ot.TrainContext(store, torch.float64)
    x = torch.from_numpy(GOLD["gd_second_x"]).double()
    y = torch.from_numpy(GOLD["gd_second_y"]).double()
    g = y * 0.5 + 0.1
    disc, _ = o8.growing_disc(y, x, 2.4, ctx, cfg)
    gen, _ = o8.growing_disc(g, x, 2.4, ctx, cfg)
    lf = torch.tensor([[0.3], [0.8]], dtype=torch.float64)
    L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc(t, x, 2.4, ctx, cfg)[0], y, g, lf, image_side=S)
    assert L["grad_norms"].shape == (2, S) and float(L["grad_penalty"]) > 0
    y2 = torch.stack([y, g], dim=2).reshape(2, S * S * 2)
    x_in, y_in = o8.refine_input(x, y2, cfg)
    assert torch.equal(y_in, y) and x_in.shape == (2, S * S * (cfg.n_inputChannels + 1))
    xi = x_in.reshape(2, S, S, -1)
    assert torch.equal(xi[..., 0], g.reshape(2, S, S))
    lo = x.reshape(2, cfg.tileSizeLow, cfg.tileSizeLow, -1)
    assert torch.equal(xi[:, 5, 9, 1:], lo[:, 5 * cfg.tileSizeLow // S, 9 * cfg.tileSizeLow // S, :])


@pytest.mark.parametrize("tag", ["gt_first", "gt_second"])
def test_temporal_critic_oracle_reproduces_the_reference_code(tag):
    """growing_disc_tempo (:868-923) on three-frame samples [B, S*S, 3]: logits and requested variables against the
    reference's own function on the TF1 shim; its gradient penalty has one norm per (sample, frame) (:1284)."""
    c, cfg = _cfg(tag)
    store = og.VarStore(seed=c["seed"])
    fr = torch.from_numpy(GOLD[tag + "_frames"]).double()
    for k, pct in enumerate(c["percentages"]):
        logits = o8.growing_disc_tempo(fr, pct, og.Context(store, torch.float64), cfg)
        ref = GOLD["%s_p%d_logits" % (tag, k)]
        assert np.abs(logits.numpy() - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), (tag, pct)
    want = {k: tuple(v) for k, v in json.loads(str(GOLD[tag + "_vars"]))}
    assert {k: tuple(v.shape) for k, v in store.values.items()} == want
    ctx = ot.TrainContext(store, torch.float64)
    rows = fr.reshape(2, -1)
    fake = rows * 0.5 + 0.1
    disc = o8.growing_disc_tempo(rows, 2.4, ctx, cfg)
    gen = o8.growing_disc_tempo(fake, 2.4, ctx, cfg)
    lf = torch.tensor([[0.3], [0.8]], dtype=torch.float64)
    L = o8.wgan_gp_losses(disc, gen, lambda t: o8.growing_disc_tempo(t, 2.4, ctx, cfg), rows, fake, lf, frames=3)
    assert L["grad_norms"].shape == (2, 3) and float(L["grad_penalty"]) > 0


def test_tensor_resample_oracle_reproduces_the_reference_code():
    """tensorResample (:545-594) executed on the shim vs the torch restatement, incl. positions that leave the tile."""
    val = torch.from_numpy(GOLD["resample_value"]).double()
    pos = torch.from_numpy(GOLD["resample_pos"]).double()
    out = o8.tensor_resample(val, pos)
    assert np.abs(out.numpy() - GOLD["resample_out"]).max() < 1e-12
    q = (pos - 0.5).floor()
    assert bool(((q < 0) | (q + 1 > 7)).any())          # the vectors do exercise out-of-range corners


def test_tempo_tiles_and_semi_lagrangian_positions_reproduce_the_reference_code():
    """getTempoinput's data side: selectRandomTempoTiles / getSemiLagrPosBatch (tools_wscale/tilecreator_t.py:1291-1413)."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tempotiles.npz"))
    xt, yt, pos = o8.tempo_tiles(G["low"], G["high"], 3, 0.5)
    assert np.array_equal(xt, G["xt"]) and np.array_equal(yt, G["yt"])
    # (the reference interpolates float32 velocities in float32 -- scipy keeps the input type -- the restatement in float64)
    assert np.abs(pos - G["pos"]).max() < 5e-6
    assert np.abs(o8.semi_lagr_positions(G["vel"][:, 0], G["dt"], 24) - G["pos_up8"]).max() < 5e-6
    assert np.abs(o8.semi_lagr_positions(G["vel"][:, 0], G["dt"], 3) - G["pos_same"]).max() < 5e-6
