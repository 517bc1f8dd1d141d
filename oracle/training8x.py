"""ORACLE (test infrastructure only): pieces of the 8x progressive-growing trainer on CPU, torch fp64.

Restates GAN/multipassGAN-8x.py for the configuration the shipped training command uses for the first network
(GAN/example_run_training.py:4: `firstNNArch 1 upsamplingMode 2 use_wgan_gp 1 batchNorm 0 gDrop 0 use_mb_stddev 0`):
  lerp :596-597, growBlockDisc :752-780, growing_disc :782-866 (spatial discriminator grown stage by stage, every stage
  blended in with lerp(old, new, percentage - (j-1))), the WGAN-GP discriminator / generator losses :1101-1143
  (tf.gradients of the critic w.r.t. the interpolated sample: a double backward here), the per-stage variable selection of
  the optimizers :1304-1362 ("%i" % 2**i in var.name), TF1 Adam (oracle.training.Adam) and the generator weight averaging
  of tf.contrib.opt.MovingAverageOptimizer(…, 0.999) :1356-1361.
Pinned by golden vectors produced by executing the reference's own growing_disc / growBlockDisc / lerp on the numpy TF1
shim (tests/golden/make_golden.py growdisc -> tests/golden/growdisc.npz); gradients are torch autograd in fp64.
Also growing_disc_tempo :868-923 (the temporal critic of three aligned frames) with its WGAN-GP loss :1262-1289, pinned the same
way (gt_first / gt_second vectors).  Out of scope: the frame alignment in front of it (advection / tensorResample, adv_flag),
loss scaling (numerically the identity), gDrop noise.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import gan as og
from . import networks as on
from . import tf_ops


def lerp(x, y, t):
    """GAN/multipassGAN-8x.py:596-597."""
    return x + (y - x) * torch.clamp(torch.as_tensor(t, dtype=x.dtype), 0.0, 1.0)


def avg_pool2(x):
    """GAN.avg_pool default (tools_wscale/GAN.py:162-169): tf.nn.avg_pool 2x2 stride 2 VALID, NHWC."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


class Cfg8x:
    """Flags of GAN/multipassGAN-8x.py that shape growing_disc (defaults of the shipped first-network command)."""

    def __init__(self, tileSizeLow=16, upRes=8, n_inputChannels=6, start_fms=256, max_fms=256, filterSize=3,
                 first_nn_arch=True, upsampleMode=1, upsampling_mode=2):
        self.upsampling_mode = int(upsampling_mode)  # 2: first network (grows in resolution); 1 / 3: refinement networks
        assert self.upsampling_mode in (1, 2, 3)
        self.tileSizeLow, self.upRes = int(tileSizeLow), int(upRes)
        self.tileSizeHigh = self.tileSizeLow * self.upRes
        self.n_inputChannels = int(n_inputChannels)
        self.start_fms, self.max_fms, self.filterSize = int(start_fms), int(max_fms), int(filterSize)
        self.first_nn_arch = bool(first_nn_arch)
        self.upsampleMode = int(upsampleMode)
        self.stages = int(round(math.log(self.upRes, 2)))


def grow_block_disc(gan, inp, upres, fms, cfg, name="d"):
    """growBlockDisc :752-780 (no batch norm, no gDrop). Returns (pooled x2 [upsampling_mode 2] or x2 [1 / 3], x1, x2)."""
    with gan.ctx.variable_scope(name + "Block%d" % upres):
        filt = [4, 4] if cfg.first_nn_arch else [cfg.filterSize, cfg.filterSize]
        out2 = min(min(fms * 2, cfg.max_fms), cfg.start_fms // 2)
        if cfg.first_nn_arch:
            c1 = fms * 3 if upres == 2 else fms * 2
            x1, _ = gan.convolutional_layer(c1, filt, og.lrelu, stride=[1], name="%s_cA%d" % (name, upres), in_layer=inp,
                                            in_channels=fms)
            x2, _ = gan.convolutional_layer(out2, filt, og.lrelu, stride=[1], name="%s_cB%d" % (name, upres), in_layer=x1)
        else:
            x1, _ = gan.convolutional_layer(fms, filt, og.lrelu, stride=[1], name="%s_cA%d" % (name, upres), in_layer=inp,
                                            in_channels=fms)
            x2, _ = gan.convolutional_layer(out2, filt, og.lrelu, stride=[1], name="%s_cB%d" % (name, upres), in_layer=x1,
                                            in_channels=fms)
        if cfg.upsampling_mode == 2:
            gan.layer = avg_pool2(gan.layer)  # outp = gan.avg_pool() :771-772
            return gan.layer, x1, x2
        return x2, x1, x2                     # :773-774


def growing_disc(in_high, in_low, percentage, ctx, cfg):
    """growing_disc :782-866. in_high [B, S*S], in_low [B, L*L*C] flat rows -> (logits [B,1], feature layers)."""
    L, S, C, u = cfg.tileSizeLow, cfg.tileSizeHigh, cfg.n_inputChannels, cfg.upRes
    with ctx.variable_scope("spatial-disc"):
        hi = in_high.reshape(-1, S, S, 1)
        lo = in_low.reshape(-1, L, L, C)[..., 0:1]                                     # :796
        lo = og.GAN(lo, ctx).avg_depool(scale=[u], mode=cfg.upsampleMode)              # :797
        hi = torch.cat([lo, hi], dim=3)                                                # :813
        feats = []
        gan = og.GAN(hi, ctx)
        x_, _ = gan.convolutional_layer(int(cfg.start_fms / u), [1, 1], None, in_layer=hi, stride=[1],
                                        name="d_cfromDensity%d" % u)                   # :818
        feats.append(lerp(torch.zeros_like(x_), x_, percentage - (cfg.stages - 1)))
        in_high_ = hi
        gan2 = og.GAN(in_high_, ctx)
        for j in range(cfg.stages, 0, -1):
            num_fms = int(min(cfg.start_fms / (2 ** j), cfg.max_fms))
            if cfg.upsampling_mode == 2:
                in_high_ = avg_pool2(in_high_)                                         # :828-829
            x_, x1, x2 = grow_block_disc(gan, x_, int(2 ** j), num_fms, cfg)           # :833
            from_dens = min(min(num_fms * 2, cfg.max_fms), cfg.start_fms // 2)
            old, _ = gan2.convolutional_layer(from_dens, [1, 1], None, stride=[1], name="d_cfromDensity%d" % (2 ** (j - 1)),
                                              in_layer=in_high_)                       # :836
            with ctx.variable_scope("blend%i" % j):
                x_ = lerp(old, x_, percentage - (j - 1))                               # :838-839
            feats.append(lerp(torch.zeros_like(x1), x1, percentage - (j - 1)))
            feats.append(lerp(torch.zeros_like(x2), x2, percentage - (j - 1)))
        if not cfg.first_nn_arch:
            f = [cfg.filterSize, cfg.filterSize]
            x1, _ = gan.convolutional_layer(32, f, og.lrelu, stride=[1], name="d_cA1", in_layer=x_)
            gan.convolutional_layer(4, f, None, stride=[1], name="d_cB1", in_layer=x1)
        else:
            # quirk reproduced: nothing moves the GAN cursor here, so flatten() below acts on gan.layer = the pooled output of
            # the LAST growBlockDisc -- the blend with d_cfromDensity1 only reaches the feature list, not the logits
            x1 = x_
        feats.append(lerp(torch.zeros_like(x1), x1, percentage))
        gan.flatten()
        gan.fully_connected_layer(1, None, name="d_l61", gain=1)                        # :862
        return gan.y(), feats


def semi_lagr_positions(vel, dt, out_side):
    """getSemiLagrPosBatch (tools_wscale/tilecreator_t.py:1341-1378) for 2-D tiles: vel [n, L, L, 3] (vx, vy, vz per low-res
    cell), dt [n] -> positions [n, S, S, 2] (y, x; cell centres at i + 0.5) = centre - centred velocity * dt, where the
    velocity is first interpolated to the S x S grid (gridInterpolBatch :1291-1314: scipy map_coordinates order 1, mode
    'nearest', at index (i + 0.5) * L / S), then centred like a MAC grid (getMACGridCenteredBatch :1319-1338: the mean of a
    component and its +1 neighbour along its own axis, the last cell repeated) and scaled by S / L."""
    vel = np.asarray(vel, np.float64)
    n, L = vel.shape[0], vel.shape[1]
    S = int(out_side)
    dt = np.asarray(dt, np.float64).reshape(n, 1, 1)
    if S == L:
        vy, vx, scale = vel[..., 1], vel[..., 0], 1.0
    else:
        c = np.clip((np.arange(S) + 0.5) * (L / S), 0.0, L - 1.0)
        lo = np.minimum(np.floor(c).astype(int), L - 1)
        hi = np.minimum(lo + 1, L - 1)
        t = c - lo

        def up(a):
            rows = a[:, lo] * (1 - t)[None, :, None] + a[:, hi] * t[None, :, None]
            return rows[:, :, lo] * (1 - t)[None, None, :] + rows[:, :, hi] * t[None, None, :]
        vy, vx, scale = up(vel[..., 1]), up(vel[..., 0]), S / L
    nxt = np.minimum(np.arange(S) + 1, S - 1)
    cy = 0.5 * (vy + vy[:, nxt, :]) * scale
    cx = 0.5 * (vx + vx[:, :, nxt]) * scale
    yy, xx = np.meshgrid(np.arange(S) + 0.5, np.arange(S) + 0.5, indexing="ij")
    return np.stack([yy[None] - cy * dt, xx[None] - cx * dt], axis=-1)


def tempo_tiles(low, high, n_t=3, dt=0.5, vel_channels=(1, 2, 3)):
    """selectRandomTempoTiles (tilecreator_t.py:1382-1413) after the tile selection: low [B, 1, T, T, C*n_t], high
    [B, 1, S, S, n_t] three-frame tiles (frames stored as channel groups) -> rows ordered (sample, frame):
    x_t [B*n_t, T*T*C], y_t [B*n_t, S*S], positions [B*n_t, S*S*2] with dt * (n_t//2, ..., -(n_t//2)) per frame."""
    low, high = np.asarray(low), np.asarray(high)
    B, _, T, _, cc = low.shape
    S, C = high.shape[2], cc // n_t
    x = low.reshape(B, 1, T, T, n_t, C).transpose(0, 4, 1, 2, 3, 5).reshape(B * n_t, T, T, C)
    y = high.reshape(B, 1, S, S, n_t, -1).transpose(0, 4, 1, 2, 3, 5).reshape(B * n_t, -1)
    dts = np.array([i * dt for i in range(n_t // 2, -n_t // 2, -1)] * B, np.float32)
    pos = semi_lagr_positions(x[..., list(vel_channels)], dts, S)
    return x.reshape(B * n_t, -1), y, pos.reshape(B * n_t, -1)


def tensor_resample(value, pos):
    """tensorResample :545-594 for 2-D data: value [B, H, W, C] sampled at pos [B, H, W, 2] (pos[..., 0] along H, pos[..., 1]
    along W, cell centres at i + 0.5): bilinear weights 1 - |pos - 0.5 - index| over floor / floor + 1, indices are NOT clamped
    (the `if 0:` block), out-of-range corners contribute 0 (what tf.gather_nd does on the GPU).  Differentiable in `value`."""
    B, H, W, C = value.shape
    q = pos - 0.5
    f = torch.floor(q).long()
    out = torch.zeros(pos.shape[:-1] + (C,), dtype=value.dtype)
    bidx = torch.arange(B).view(B, 1, 1).expand(pos.shape[:-1])
    for c0 in (0, 1):
        for c1 in (0, 1):
            i0, i1 = f[..., 0] + c0, f[..., 1] + c1
            w = (1.0 - (q[..., 0] - i0.to(q.dtype)).abs()) * (1.0 - (q[..., 1] - i1.to(q.dtype)).abs())
            ok = (i0 >= 0) & (i0 < H) & (i1 >= 0) & (i1 < W)
            v = value[bidx, i0.clamp(0, H - 1), i1.clamp(0, W - 1)]
            out = out + v * (w * ok.to(q.dtype)).unsqueeze(-1)
    return out


def growing_disc_tempo(frames, percentage, ctx, cfg):
    """growing_disc_tempo :868-923 (useVelInTDisc 0, no batch norm / gDrop / minibatch stddev): the unconditional critic of
    three aligned frames. frames [B, S*S, 3] (what `tf.transpose(reshape(., [-1, 3, n_output]), [0, 2, 1])` of :1213-1214
    hands over) -> logits [B, 1].  Same growing structure as growing_disc with the name prefix "t" and a 3-channel image."""
    S, u = cfg.tileSizeHigh, cfg.upRes
    with ctx.variable_scope("tempo-disc"):
        img = frames.reshape(-1, S, S, 3)                                              # :879
        gan = og.GAN(img, ctx)
        x, _ = gan.convolutional_layer(int(cfg.start_fms / u), [1, 1], None, in_layer=img, stride=[1],
                                       name="t_cfromDensity%d" % u)                    # :882
        in_high = img
        gan2 = og.GAN(in_high, ctx)
        for j in range(cfg.stages, 0, -1):
            num_fms = int(min(cfg.start_fms / (2 ** j), cfg.max_fms))
            if cfg.upsampling_mode == 2:
                in_high = avg_pool2(in_high)                                           # :889-890
            x, _, _ = grow_block_disc(gan, x, int(2 ** j), num_fms, cfg, name="t")     # :894
            from_dens = min(min(num_fms * 2, cfg.max_fms), cfg.start_fms // 2)
            old, _ = gan2.convolutional_layer(from_dens, [1, 1], None, stride=[1], name="t_cfromDensity%d" % (2 ** (j - 1)),
                                              in_layer=in_high)                        # :896
            with ctx.variable_scope("blend%i" % j):
                x = lerp(old, x, percentage - (j - 1))                                 # :899-904
        if not cfg.first_nn_arch:
            f = [cfg.filterSize, cfg.filterSize]
            x1, _ = gan.convolutional_layer(32, f, og.lrelu, stride=[1], name="t_cA1", in_layer=x)
            gan.convolutional_layer(4, f, None, stride=[1], name="t_cB1", in_layer=x1)
        # (firstNNArch: same cursor quirk as growing_disc -- flatten() sees the pooled output of the last growBlockDisc)
        gan.flatten()
        gan.fully_connected_layer(1, None, name="t_l61", gain=1)                        # :919
        return gan.y()


def growing_gen_train(x_rows, percentage, ctx, cfg, pixel_norm=True, addBicubicUpsample=True):
    """growing_gen with output=False (GAN/multipassGAN-8x.py:700-750): every stage emits a density (1x1 conv, gain 1) plus the
    residual input density, blended with the density of the previous stage by lerp(old, new, percentage - (j-1)).
    upsampling_mode 2 (firstNNArch): x_rows [B, L*L*C], nearest x2 per stage, bicubic residual; upsampling_mode 1 / 3
    (refinement networks): x_rows [B, S*S*(C+1)] = concat(first-pass density, resized low-res fields), two head resBlocks,
    no resampling, the residual is channel 0 of the input.  Returns the flat rows [B, S*S]."""
    ocfg = on.make_cfg_out(cfg.tileSizeLow, cfg.upRes, cfg.n_inputChannels, pixel_norm=pixel_norm,
                           addBicubicUpsample=addBicubicUpsample, upsampleMode=cfg.upsampleMode)
    L, S, C = cfg.tileSizeLow, cfg.tileSizeHigh, cfg.n_inputChannels
    first = cfg.upsampling_mode == 2
    assert first == cfg.first_nn_arch, "built: firstNNArch with upsampling_mode 2, the plain res-net with modes 1 / 3"
    with ctx.variable_scope("generator"):
        _in = x_rows.reshape(-1, L, L, C) if first else x_rows.reshape(-1, S, S, C + 1)   # :702-705
        gan = og.GAN(_in, ctx)
        if first:
            x_g = _in                                                                   # first_nn_arch :712-713
        else:
            m = min(cfg.max_fms, cfg.start_fms // 2)
            x_g = on._resblock_out(gan, ocfg, _in, 16, m // 8, False, "1", cfg.filterSize, False)        # :715
            x_g = on._resblock_out(gan, ocfg, x_g, m // 4, m // 2, False, "2", cfg.filterSize, False)    # :716
        old, _ = og.GAN(x_g, ctx).convolutional_layer(1, [1, 1], None, stride=[1], name="g_cdensOut1", in_layer=x_g, gain=1)
        for j in range(1, cfg.stages + 1):
            num_fms = min(int(cfg.start_fms / (2 ** j)), cfg.max_fms)
            x_g, dens = on._grow_block_gen(gan, ctx, ocfg, x_g, int(2 ** j), num_fms, False, False, False, first,
                                           cfg.filterSize, first, True)
            if addBicubicUpsample:                                                      # :735-739
                if first:
                    dens = dens + og.GAN(_in[..., 0:1], ctx).avg_depool(mode=2, scale=[int(2 ** j)])
                else:
                    dens = dens + _in[..., 0:1]
            with ctx.variable_scope("growingPart%i" % j):
                if first:
                    old = og.GAN(old, ctx).avg_depool(mode=1)                           # :744
                old = lerp(old, dens, percentage - (j - 1))                             # :748-750
        return old.reshape(-1, S * S)


def refine_input(x_rows, y_rows, cfg):
    """x_in / y_in of the refinement networks' training graph (:1042-1044, 1061-1062): the target rows carry two channels,
    y[..., 0] = the high-res density to learn, y[..., 1] = the first-pass density; the network input is
    concat(y[..., 1], nearest resize of the low-res fields).  Returns (x_in rows [B, S*S*(C+1)], y_in [B, S*S])."""
    L, S, C = cfg.tileSizeLow, cfg.tileSizeHigh, cfg.n_inputChannels
    y2 = y_rows.reshape(-1, S, S, 2)
    lo = tf_ops.resize_nearest(x_rows.reshape(-1, L, L, C), S, S)
    x_in = torch.cat([y2[..., 1:2], lo], dim=3)
    return x_in.reshape(-1, S * S * (C + 1)), y2[..., 0].reshape(-1, S * S)


def wgan_gp_losses(disc, gen, d_out_fn, y_in, gen_y, lerp_factor, wgan_lambda=10.0, wgan_target=1.0, wgan_epsilon=0.001,
                   weight_dld=1.0, image_side=None, frames=None):
    """Discriminator / generator critic losses with use_wgan_gp (:1101-1143). d_out_fn(y) -> critic logits.
    lerp_factor: the tf.random_uniform([B, 1]) sample of :1120 (fed in, so that both sides use the same numbers).
    image_side: upsampling_mode 1 / 3 keeps the samples as [B, S, S, 1] images there (:1047, 1062), so the reference's
    `reduce_sum(..., axis=1)` (:1130) sums over the image ROWS only: one gradient norm per (sample, column).  Reproduced:
    pass S to get that reduction on the flat rows.
    frames: the temporal critic's samples are [B, S*S, frames] (:1213-1214, 1279-1285), so the same reduce_sum(axis=1) gives
    one norm per (sample, frame): pass 3 with rows of S*S*3 values."""
    d_loss = (-disc).mean() * weight_dld + gen.mean()
    y_gp = (lerp_factor * y_in + (1 - lerp_factor) * gen_y).detach().requires_grad_(True)
    d_out_loss = d_out_fn(y_gp).mean()
    grads = torch.autograd.grad(d_out_loss, y_gp, create_graph=True)[0]
    if frames is not None:
        norms = torch.sqrt(((grads.reshape(grads.shape[0], -1, frames) + 1e-4) ** 2).sum(dim=1))   # :1284
    elif image_side is None:
        norms = torch.sqrt(((grads + 1e-4) ** 2).sum(dim=1))                           # :1130
    else:
        norms = torch.sqrt(((grads.reshape(-1, image_side, image_side) + 1e-4) ** 2).sum(dim=1))
    gp = (wgan_lambda * (norms - wgan_target) ** 2).mean()
    eps_pen = (disc ** 2).mean()
    return dict(disc_loss=d_loss + eps_pen * wgan_epsilon + gp, grad_penalty=gp, epsilon_penalty=eps_pen,
                g_loss_d=(-gen).mean(), grad_norms=norms.detach())


def stage_variables(names, z, n_stages=3):
    """Variable subset optimizer z updates (:1332-1338 / :1349-1355): all of them at the last stage, else those whose name
    contains "%i" % 2**i for an i in [0, z+1] -- a SUBSTRING test, so "1" also selects every name with a 1 in it
    (d_cfromDensity16 would match, d_l61 does; reproduced as is)."""
    if z == n_stages - 1:
        return list(names)
    out = []
    for i in range(0, z + 2):
        out.extend(n for n in names if ("%i" % (2 ** i)) in n and n not in out)
    return out


def ema_init(values):
    """ExponentialMovingAverage shadows start at the variables' initial values."""
    return {k: np.array(v, dtype=np.float64) for k, v in values.items()}


def ema_update(shadow, values, decay=0.999):
    """tf.contrib.opt.MovingAverageOptimizer(opt, 0.999) = ExponentialMovingAverage without num_updates, applied after every
    optimizer step to the variables that step updates: s -= (1 - decay) * (s - v)."""
    for k, v in values.items():
        shadow[k] = shadow[k] - (1.0 - decay) * (shadow[k] - np.asarray(v, np.float64))
    return shadow
