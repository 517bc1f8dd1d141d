"""ORACLE (test infrastructure only - never imported by the product path).

Restatement of the reference network builders on top of oracle.gan.GAN:
  gen_resnet      GAN/multipassGAN-4x.py:505-569   (the "multipassGAN-4x" generator)
  disc_binclass   GAN/multipassGAN-4x.py:572-620   (4x spatial discriminator)
  growing_gen     GAN/multipassGAN-out.py:220-338  (8x progressive-growing generators, apply mode)
Script-level globals of the reference (upRes, tileSizeLow, ...) are carried in a `cfg` namespace.
PARITY UNPINNED.
"""
import math
from types import SimpleNamespace

import torch

from . import tf_ops
from .gan import GAN, lrelu, relu


def make_cfg_4x(tileSizeLow, upRes=4, n_inputChannels=4, upsampling_mode=2, batch_norm=True, bn_decay=0.999):
    return SimpleNamespace(tileSizeLow=tileSizeLow, upRes=upRes, tileSizeHigh=tileSizeLow * upRes,
                           n_inputChannels=n_inputChannels, upsampling_mode=upsampling_mode,
                           batch_norm=batch_norm, bn_decay=bn_decay)


# ---------------------------------------------------------------- 4x generator
def _resblock_4x(gan, inp, s1, s2, use_batch_norm, train, rbId, filter_size):
    """GAN/multipassGAN-4x.py:505-526. Conv B takes its input from the cursor (App. D.2)."""
    f = [filter_size, filter_size]
    gan.convolutional_layer(s1, f, relu, stride=[1], name="g_cA%d" % rbId, in_layer=inp,
                            batch_norm=use_batch_norm, train=train)
    gc2, _ = gan.convolutional_layer(s2, f, None, stride=[1], name="g_cB%d" % rbId,
                                     batch_norm=use_batch_norm, train=train)
    gs1, _ = gan.convolutional_layer(s2, [1, 1], None, stride=[1], name="g_s%d" % rbId, in_layer=inp,
                                     batch_norm=use_batch_norm, train=train)
    return torch.relu(gc2 + gs1)


def gen_resnet(_in, ctx, cfg, train=False):
    """GAN/multipassGAN-4x.py:528-569. _in: flat [N, n_input]; returns flat [N, tileSizeHigh^2]."""
    with ctx.variable_scope("generator"):
        C = cfg.n_inputChannels
        if cfg.upsampling_mode == 2:
            x = _in.reshape(-1, cfg.tileSizeLow, cfg.tileSizeLow, C)
        elif cfg.upsampling_mode in (1, 3):
            x = _in.reshape(-1, cfg.tileSizeHigh, cfg.tileSizeHigh, C)
        else:
            raise NotImplementedError("upsampling_mode 0 is not on the benchmarked path")
        filterSize = 5
        gan = GAN(x, ctx)
        if cfg.upsampling_mode == 2:
            inp = gan.max_depool(height_factor=cfg.upRes, width_factor=cfg.upRes)
        else:
            inp = x
        bn = cfg.batch_norm
        ru1 = _resblock_4x(gan, inp, C * 2, C * 8, bn, train, 0, filterSize)
        ru2 = _resblock_4x(gan, ru1, 128, 128, bn, train, 1, filterSize)
        ru3 = _resblock_4x(gan, ru2, 32, 8, bn, train, 2, filterSize)
        ru4 = _resblock_4x(gan, ru3, 2, 1, False, train, 3, filterSize)
        return ru4.reshape(-1, cfg.tileSizeHigh * cfg.tileSizeHigh), gan


# ---------------------------------------------------------------- 4x discriminator
def disc_binclass(in_low, in_high, ctx, cfg, train=False, use_batch_norm=True):
    """GAN/multipassGAN-4x.py:572-620 (2-D branch). Returns (logits, d1, d2, d3, d4)."""
    with ctx.variable_scope("discriminator"):
        n_input = in_low.shape[1]
        # App. D.5: the first n_input/C floats of the flat INTERLEAVED row
        in_low = in_low[:, : n_input // cfg.n_inputChannels]
        if cfg.upsampling_mode == 2:
            lo = in_low.reshape(-1, cfg.tileSizeLow, cfg.tileSizeLow, 1)
            lo = tf_ops.resize_nearest(lo, cfg.tileSizeHigh, cfg.tileSizeHigh)
        else:
            lo = in_low.reshape(-1, cfg.tileSizeHigh, cfg.tileSizeHigh, 1)
        hi = in_high.reshape(-1, cfg.tileSizeHigh, cfg.tileSizeHigh, 1)
        gan = GAN(torch.cat([lo, hi], dim=-1), ctx, bn_decay=cfg.bn_decay)
        f = [4, 4]
        d1, _ = gan.convolutional_layer(32, f, lrelu, stride=[2], name="d_c1")
        d2, _ = gan.convolutional_layer(64, f, lrelu, stride=[2], name="d_c2", batch_norm=use_batch_norm, train=train)
        d3, _ = gan.convolutional_layer(128, f, lrelu, stride=[2], name="d_c3", batch_norm=use_batch_norm, train=train)
        d4, _ = gan.convolutional_layer(256, f, lrelu, stride=[1], name="d_c4", batch_norm=use_batch_norm, train=train)
        gan.flatten()
        gan.fully_connected_layer(1, None, name="d_l5")
        return gan.y(), d1, d2, d3, d4


# ---------------------------------------------------------------- out.py growing generators
def make_cfg_out(tileSizeLow, upRes=8, n_inputChannels=4, pixel_norm=True, batch_norm=False, upsampleMode=1,
                 addBicubicUpsample=True, usePixelShuffle=False):
    return SimpleNamespace(tileSizeLow=tileSizeLow, upRes=upRes, tileSizeHigh=tileSizeLow * upRes,
                           n_inputChannels=n_inputChannels, pixel_norm=pixel_norm, batch_norm=batch_norm,
                           upsampleMode=upsampleMode, addBicubicUpsample=addBicubicUpsample,
                           usePixelShuffle=usePixelShuffle)


def _resblock_out(gan, cfg, inp, s1, s2, use_batch_norm, name, filter_size, train):
    """GAN/multipassGAN-out.py:220-237."""
    f = [filter_size, filter_size]
    gc1, _ = gan.convolutional_layer(s1, f, relu, stride=[1], name="g_cA_" + name, in_layer=inp,
                                     batch_norm=use_batch_norm, train=train)
    if cfg.pixel_norm:
        gc1 = gan.pixel_norm(gc1)
    gc2, _ = gan.convolutional_layer(s2, f, None, stride=[1], name="g_cB_" + name,
                                     batch_norm=use_batch_norm, train=train)
    gs1, _ = gan.convolutional_layer(s2, [1, 1], None, stride=[1], name="g_s_" + name, in_layer=inp,
                                     batch_norm=use_batch_norm, train=train)
    res = torch.relu(gc2 + gs1)
    if cfg.pixel_norm:
        res = gan.pixel_norm(res)
    return res


def _grow_block_gen(gan, ctx, cfg, inp, upres, fms, use_batch_norm, train, output, firstGen, filterSize,
                    first_nn_arch, use_res_net):
    """GAN/multipassGAN-out.py:239-284."""
    with ctx.variable_scope("genBlock%d" % upres):
        if firstGen:
            assert not cfg.usePixelShuffle
            inDepool = gan.avg_depool(mode=cfg.upsampleMode)  # acts on gan.layer (App. D.1)
        else:
            inDepool = inp
        fs = filterSize
        if first_nn_arch:
            if upres == 2:
                outp = _resblock_out(gan, cfg, inDepool, fms, fms, use_batch_norm, "first", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "second", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "third", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "fourth", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "fifth", fs, train)
            elif upres == 4:
                outp = _resblock_out(gan, cfg, inDepool, fms * 2, fms, use_batch_norm, "first", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "second", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "third", fs, train)
            if upres == 8:
                outp = _resblock_out(gan, cfg, inDepool, fms * 2, fms, use_batch_norm, "first", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms, fms, use_batch_norm, "second", fs, train)
        else:
            if use_res_net:
                outp = _resblock_out(gan, cfg, inDepool, fms, fms, use_batch_norm, "first", fs, train)
                outp = _resblock_out(gan, cfg, outp, fms // 2, fms // 2, use_batch_norm, "second", fs, train)
            else:
                i2, _ = gan.convolutional_layer(fms, [fs, fs], lrelu, stride=[1], name="g_cA%d" % upres,
                                                in_layer=inDepool, batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    i2 = gan.pixel_norm(i2)
                outp, _ = gan.convolutional_layer(fms, [fs, fs], lrelu, stride=[1], name="g_cB%d" % upres,
                                                  in_layer=i2, batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    outp = gan.pixel_norm(outp)
        if not output:
            outpDens, _ = GAN(outp, ctx, bn_decay=0.0).convolutional_layer(
                1, [1, 1], None, stride=[1], name="g_cdensOut%d" % upres, in_layer=outp, batch_norm=False,
                train=train, gain=1)
            return outp, outpDens
        return outp


def growing_gen(_in, ctx, cfg, use_batch_norm=False, train=False, currentUpres=3, output=True, firstGen=True,
                filterSize=3, startFms=256, maxFms=256, add_adj_idcs=False, first_nn_arch=False,
                use_res_net=True):
    """GAN/multipassGAN-out.py:286-338. _in is NHWC-reshapeable: firstGen -> [N, L*L*(C[+2])],
    else [N, S, S, C+1] (already concatenated by the sampler wiring :357)."""
    with ctx.variable_scope("generator"):
        n_channels = cfg.n_inputChannels + (2 if add_adj_idcs else 0)
        if firstGen:
            x = _in.reshape(-1, cfg.tileSizeLow, cfg.tileSizeLow, n_channels)
        else:
            x = _in.reshape(-1, cfg.tileSizeHigh, cfg.tileSizeHigh, n_channels + 1)
        gan = GAN(x, ctx, bn_decay=0.0)
        fs = filterSize
        if first_nn_arch:
            x_g = x
        else:
            if use_res_net:
                m = min(maxFms, startFms // 2)
                x_g = _resblock_out(gan, cfg, x, 16, m // 8, False, "1", fs, train)
                x_g = _resblock_out(gan, cfg, x_g, m // 4, m // 2, False, "2", fs, train)
            else:
                x_g, _ = gan.convolutional_layer(32, [fs, fs], lrelu, stride=[1], name="g_cA%d" % 1, in_layer=x,
                                                 batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    x_g = gan.pixel_norm(x_g)
                x_g, _ = gan.convolutional_layer(min(startFms // 2, maxFms), [fs, fs], lrelu, stride=[1],
                                                 name="g_cB%d" % 1, in_layer=x_g, batch_norm=use_batch_norm,
                                                 train=train)
                if cfg.pixel_norm:
                    x_g = gan.pixel_norm(x_g)
        _dens = None
        for j in range(1, currentUpres + 1):
            num_fms = min(int(startFms / (2 ** j)), maxFms)
            if (not output) or j == currentUpres:
                x_g, _dens = _grow_block_gen(gan, ctx, cfg, x_g, int(2 ** j), num_fms, use_batch_norm, train, False,
                                             firstGen, fs, first_nn_arch, use_res_net)
            else:
                x_g = _grow_block_gen(gan, ctx, cfg, x_g, int(2 ** j), num_fms, use_batch_norm, train, output,
                                      firstGen, fs, first_nn_arch, use_res_net)
            if cfg.addBicubicUpsample and j == currentUpres:
                if firstGen:
                    dens_in = x[:, :, :, 0:1]
                    _dens = _dens + GAN(dens_in, ctx).avg_depool(mode=2, scale=[int(2 ** j)])
                else:
                    _dens = _dens + x[:, :, :, 0:1]
        return _dens.reshape(-1, cfg.tileSizeHigh * cfg.tileSizeHigh), gan


def sampler_input_2(x_flat, y_flat, cfg):
    """GAN/multipassGAN-out.py:357: concat(first-pass density, nearest-resized low-res fields)."""
    lo = x_flat.reshape(-1, cfg.tileSizeLow, cfg.tileSizeLow, cfg.n_inputChannels)
    lo_up = tf_ops.resize_nearest(lo, cfg.tileSizeHigh, cfg.tileSizeHigh)
    hi = y_flat.reshape(-1, cfg.tileSizeHigh, cfg.tileSizeHigh, 1)
    return torch.cat([hi, lo_up], dim=3)


def log2i(u):
    return int(round(math.log(u, 2)))
