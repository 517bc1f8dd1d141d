"""Generator graph builders, written against the drop-in layer API (GAN.py) exactly the way the
reference scripts write them against tools_wscale/GAN.py:

  resBlock / gen_resnet          GAN/multipassGAN-4x.py:505-569   (the "multipassGAN-4x" generator)
  resBlock_out / growBlockGen /
  growing_gen                    GAN/multipassGAN-out.py:220-338  (8x progressive-growing generators)
  sampler_2_input                GAN/multipassGAN-out.py:357      (second/third generator input wiring)

The scripts' module-level globals (tileSizeLow, upRes, pixel_norm, ...) are carried by a config
object instead of Python globals so several networks can coexist in one process.
"""
import math
from types import SimpleNamespace

from . import graph as G
from .GAN import GAN, lrelu, relu


# ====================================================================== 4x (multipassGAN-4x.py)
def config_4x(tileSizeLow, upRes=4, n_inputChannels=4, upsampling_mode=2, batch_norm=True, bn_decay=0.999):
    """Flags of GAN/multipassGAN-4x.py:32-144 that shape gen_resnet (batchNorm defaults to True, :85)."""
    return SimpleNamespace(tileSizeLow=tileSizeLow, upRes=upRes, tileSizeHigh=tileSizeLow * upRes,
                           n_inputChannels=n_inputChannels, upsampling_mode=upsampling_mode,
                           batch_norm=batch_norm, bn_decay=bn_decay, rbId=0)


def resBlock(gan, cfg, inp, s1, s2, reuse, use_batch_norm, filter_size=3, train=False):
    """GAN/multipassGAN-4x.py:505-526 (2-D branch). Conv B reads the cursor (App. D.2)."""
    filter = [filter_size, filter_size]
    filter1 = [1, 1]
    rbId = cfg.rbId
    gc1, _ = gan.convolutional_layer(s1, filter, relu, stride=[1], name="g_cA%d" % rbId, in_layer=inp, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    gc2, _ = gan.convolutional_layer(s2, filter, None, stride=[1], name="g_cB%d" % rbId, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    gs1, _ = gan.convolutional_layer(s2, filter1, None, stride=[1], name="g_s%d" % rbId, in_layer=inp, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    resUnit1 = G.relu(G.add(gc2, gs1))
    cfg.rbId += 1
    return resUnit1


def gen_resnet(_in, cfg, reuse=False, use_batch_norm=None, train=False):
    """GAN/multipassGAN-4x.py:528-569. `_in`: placeholder [None, n_input]; returns [None, tileSizeHigh^2]."""
    if use_batch_norm is None:
        use_batch_norm = cfg.batch_norm
    with G.variable_scope("generator", reuse=reuse):
        C = cfg.n_inputChannels
        if cfg.upsampling_mode == 2:
            _in = G.reshape(_in, shape=[-1, cfg.tileSizeLow, cfg.tileSizeLow, C])
        elif cfg.upsampling_mode in (1, 3):
            _in = G.reshape(_in, shape=[-1, cfg.tileSizeHigh, cfg.tileSizeHigh, C])
        else:
            raise NotImplementedError("upsamplingMode 0 is outside the benchmarked path")
        cfg.rbId = 0
        filterSize = 5
        gan = GAN(_in)
        if cfg.upsampling_mode == 2:
            inp = gan.max_depool(height_factor=cfg.upRes, width_factor=cfg.upRes)
        else:
            inp = _in
        ru1 = resBlock(gan, cfg, inp, C * 2, C * 8, reuse, use_batch_norm, filterSize, train)
        ru2 = resBlock(gan, cfg, ru1, 128, 128, reuse, use_batch_norm, filterSize, train)
        inRu3 = ru2
        ru3 = resBlock(gan, cfg, inRu3, 32, 8, reuse, use_batch_norm, filterSize, train)
        ru4 = resBlock(gan, cfg, ru3, 2, 1, reuse, False, filterSize, train)
        resF = G.reshape(ru4, shape=[-1, cfg.tileSizeHigh * cfg.tileSizeHigh])
        cfg.DOFs = gan.getDOFs()
        return resF


# ====================================================================== 8x (multipassGAN-out.py)
def config_out(tileSizeLow, upRes=8, n_inputChannels=4, pixel_norm=True, batch_norm=False, upsampleMode=1,
               addBicubicUpsample=True, usePixelShuffle=False):
    """Flags of GAN/multipassGAN-out.py:28-99 that shape growing_gen."""
    return SimpleNamespace(tileSizeLow=tileSizeLow, upRes=upRes, tileSizeHigh=tileSizeLow * upRes,
                           n_inputChannels=n_inputChannels, pixel_norm=pixel_norm, batch_norm=batch_norm,
                           upsampleMode=upsampleMode, addBicubicUpsample=addBicubicUpsample,
                           usePixelShuffle=usePixelShuffle)


def resBlock_out(gan, cfg, inp, s1, s2, reuse, use_batch_norm, name, filter_size=3, train=False):
    """GAN/multipassGAN-out.py:220-237."""
    filter = [filter_size, filter_size]
    filter1 = [1, 1]
    gc1, _ = gan.convolutional_layer(s1, filter, relu, stride=[1], name="g_cA_" + name, in_layer=inp, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    if cfg.pixel_norm:
        gc1 = gan.pixel_norm(gc1)
    gc2, _ = gan.convolutional_layer(s2, filter, None, stride=[1], name="g_cB_" + name, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    gs1, _ = gan.convolutional_layer(s2, filter1, None, stride=[1], name="g_s_" + name, in_layer=inp, reuse=reuse,
                                     batch_norm=use_batch_norm, train=train)
    resUnit1 = G.relu(G.add(gc2, gs1))
    if cfg.pixel_norm:
        resUnit1 = gan.pixel_norm(resUnit1)
    return resUnit1


def growBlockGen(gan, cfg, inp, upres, fms, use_batch_norm, train, reuse, output=False, firstGen=True, filterSize=3,
                 first_nn_arch=False, use_res_net=True):
    """GAN/multipassGAN-out.py:239-284."""
    with G.variable_scope("genBlock%d" % (upres), reuse=reuse):
        if firstGen:
            if cfg.usePixelShuffle:
                raise NotImplementedError("usePixelShuffle 1 is not used by the shipped configurations")
            inDepool = gan.avg_depool(mode=cfg.upsampleMode)  # acts on gan.layer, not on `inp` (App. D.1)
        else:
            inDepool = inp
        filter = [filterSize, filterSize]
        if first_nn_arch:
            if upres == 2:
                outp = resBlock_out(gan, cfg, inDepool, fms, fms, reuse, use_batch_norm, "first", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "second", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "third", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "fourth", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "fifth", filter[0], train)
            elif upres == 4:
                outp = resBlock_out(gan, cfg, inDepool, fms * 2, fms, reuse, use_batch_norm, "first", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "second", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "third", filter[0], train)
            if upres == 8:
                outp = resBlock_out(gan, cfg, inDepool, fms * 2, fms, reuse, use_batch_norm, "first", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms, fms, reuse, use_batch_norm, "second", filter[0], train)
        else:
            if use_res_net:
                outp = resBlock_out(gan, cfg, inDepool, fms, fms, reuse, use_batch_norm, "first", filter[0], train)
                outp = resBlock_out(gan, cfg, outp, fms // 2, fms // 2, reuse, use_batch_norm, "second", filter[0],
                                    train)
            else:
                inp, _ = gan.convolutional_layer(fms, filter, lrelu, stride=[1], name="g_cA%d" % (upres),
                                                 in_layer=inDepool, reuse=reuse, batch_norm=use_batch_norm,
                                                 train=train)
                if cfg.pixel_norm:
                    inp = gan.pixel_norm(inp)
                outp, _ = gan.convolutional_layer(fms, filter, lrelu, stride=[1], name="g_cB%d" % (upres),
                                                  in_layer=inp, reuse=reuse, batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    outp = gan.pixel_norm(outp)
        if not output:
            outpDens, _ = GAN(outp, bn_decay=0.0).convolutional_layer(
                1, [1, 1], None, stride=[1], name="g_cdensOut%d" % (upres), in_layer=outp, reuse=reuse,
                batch_norm=False, train=train, gain=1)
            return outp, outpDens
        return outp


def growing_gen(_in, cfg, percentage=None, reuse=False, use_batch_norm=False, train=False, currentUpres=2,
                output=False, firstGen=True, filterSize=3, startFms=256, maxFms=256, add_adj_idcs=False,
                first_nn_arch=False, use_res_net=True):
    """GAN/multipassGAN-out.py:286-338. `percentage` is accepted and ignored like in the reference's
    output graphs (App. D.11)."""
    with G.variable_scope("generator", reuse=reuse):
        n_channels = cfg.n_inputChannels
        if add_adj_idcs:
            n_channels += 2
        if firstGen:
            _in = G.reshape(_in, shape=[-1, cfg.tileSizeLow, cfg.tileSizeLow, n_channels])
        else:
            _in = G.reshape(_in, shape=[-1, cfg.tileSizeHigh, cfg.tileSizeHigh, n_channels + 1])
        gan = GAN(_in, bn_decay=0.0)
        filter = [filterSize, filterSize]
        if first_nn_arch:
            x_g = _in
        else:
            if use_res_net:
                m = min(maxFms, startFms // 2)
                x_g = resBlock_out(gan, cfg, _in, 16, m // 8, reuse, False, "1", filter[0], train)
                x_g = resBlock_out(gan, cfg, x_g, m // 4, m // 2, reuse, False, "2", filter[0], train)
            else:
                x_g, _ = gan.convolutional_layer(32, filter, lrelu, stride=[1], name="g_cA%d" % (1), in_layer=_in,
                                                 reuse=reuse, batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    x_g = gan.pixel_norm(x_g)
                x_g, _ = gan.convolutional_layer(min(startFms // 2, maxFms), filter, lrelu, stride=[1],
                                                 name="g_cB%d" % (1), in_layer=x_g, reuse=reuse,
                                                 batch_norm=use_batch_norm, train=train)
                if cfg.pixel_norm:
                    x_g = gan.pixel_norm(x_g)
        _dens = None
        for j in range(1, currentUpres + 1):
            num_fms = min(int(startFms / (2 ** j)), maxFms)
            if not output or j == currentUpres:
                x_g, _dens = growBlockGen(gan, cfg, x_g, int(2 ** (j)), num_fms, use_batch_norm, train, reuse, False,
                                          firstGen, filterSize, first_nn_arch, use_res_net)
            else:
                x_g = growBlockGen(gan, cfg, x_g, int(2 ** (j)), num_fms, use_batch_norm, train, reuse, output,
                                   firstGen, filterSize, first_nn_arch, use_res_net)
            if cfg.addBicubicUpsample:
                if j == currentUpres:
                    if firstGen:
                        _dens = G.add(_dens, GAN(G.slice_channels(_in, 0, 1)).avg_depool(mode=2, scale=[int(2 ** (j))]))
                    else:
                        _dens = G.add(_dens, G.slice_channels(_in, 0, 1))
        resF = G.reshape(_dens, shape=[-1, cfg.tileSizeHigh * cfg.tileSizeHigh])
        cfg.DOFs = gan.getDOFs()
        return resF


def sampler_2_input(x, y, cfg):
    """GAN/multipassGAN-out.py:357 / :363: concat(first-pass density, nearest-resized low-res fields)."""
    lo = G.reshape(x, shape=[-1, cfg.tileSizeLow, cfg.tileSizeLow, cfg.n_inputChannels])
    lo_up = G.resize_images(lo, [cfg.tileSizeHigh, cfg.tileSizeHigh], method=1)
    hi = G.reshape(y, shape=[-1, cfg.tileSizeHigh, cfg.tileSizeHigh, 1])
    return G.concat((hi, lo_up), axis=3)


def log2i(u):
    return int(round(math.log(u, 2)))
