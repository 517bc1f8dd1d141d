"""CPU tests of the host side: drop-in layer API bookkeeping, graph -> fused launch list (dry engine),
weight initialisation parity with the oracle, the C-ABI surface, and flag parsing."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi, engine, graph as G, networks as N, pipeline as P, weights as W
from mpgan_b200.GAN import GAN, lrelu, relu
from oracle import gan as og
from oracle import networks as on

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _labels(net):
    return [l for l, _ in net.steps]


def test_gen_resnet_graph_matches_reference_structure():
    G.reset_default_graph()
    cfg = N.config_4x(32)
    out = N.gen_resnet(G.placeholder([None, 32 * 32 * 4], "x"), cfg)
    g = G.get_default_graph()
    assert out.shape == (None, 128 * 128)
    assert cfg.DOFs == 634214
    names = list(g.variables)
    assert names[0] == "generator/g_cA0/weight" and "generator/g_cB2/moving_variance" in names
    assert "generator/g_cA3/gamma" not in names  # ru4 built with batch norm off (GAN/multipassGAN-4x.py:564)
    w = W.init_graph_variables(g, 1)
    net = engine.CompiledNet(out, w, 8, dry=True)
    lab = _labels(net)
    # ru1 (pack + conv A + conv B/shortcut) is ONE fused thin-resblock launch reading the fp32 rows through the x4 view
    assert len(lab) == 7 and lab[0].startswith("resblock[hm] g_cA0>g_cB0+g_s0 k5/5/1 4->8->32 128x128 relu in_up4")
    assert "g_cB1+g_s1 k5/1 128/32->128" in lab[2] and lab[2].startswith("conv[tc]")
    assert abs(net.flops / (8 * 128 * 128) - 1267412) < 1e-6  # FLOP per output voxel, SURVEY App. A.1


def test_growing_gen_graphs_and_fusion():
    cfg = N.config_out(16, upRes=8)
    G.reset_default_graph()
    out = P.build_out_graph(1, P.SHIPPED_8X[1], cfg)
    w = W.init_graph_variables(G.get_default_graph(), 1)
    assert sum(v.size for v in w.values()) == 1864961
    net = engine.CompiledNet(out, w, 8, dry=True)
    lab = _labels(net)
    assert sum("up2" in l for l in lab) == 2  # genBlock2->4 and 4->8 nearest x2 fused into the producer epilogue
    assert all(" pn" in l for l in lab if l.startswith("conv") and "cdensOut" not in l)
    # g_cdensOut8 (1x1 -> one channel) and the bicubic residual are ONE launch (mpg_dens_out)
    assert lab[-1].startswith("dens_out g_cdensOut8 k1 32->1 + residual mode 2")
    assert abs(net.flops / (8 * 128 * 128) - 521216) < 1e-6
    G.reset_default_graph()
    out2 = P.build_out_graph(2, P.SHIPPED_8X[2], cfg)
    w2 = W.init_graph_variables(G.get_default_graph(), 1)
    assert sum(v.size for v in w2.values()) == 774301
    net2 = engine.CompiledNet(out2, w2, 2, dry=True)
    assert _labels(net2)[-1].startswith("dens_out g_cdensOut8 k1 12->1 + residual mode 0") and _labels(net2)[0] == "pack 128x128x5"
    assert abs(net2.flops / (2 * 128 * 128) - 1546768) < 1e-6
    assert set(net2.placeholders) == {"x", "y"}


def test_variable_names_and_values_match_oracle():
    for L, mode in ((8, 2), (4, 1)):
        G.reset_default_graph()
        cfg = N.config_4x(L, upsampling_mode=mode)
        n_in = L * L * 4 if mode == 2 else (4 * L) ** 2 * 4
        N.gen_resnet(G.placeholder([None, n_in], "x"), cfg)
        w = W.init_graph_variables(G.get_default_graph(), 9)
        store = og.VarStore(seed=9)
        on.gen_resnet(torch.zeros(1, n_in), og.Context(store, torch.float32), on.make_cfg_4x(L, upsampling_mode=mode))
        assert set(store.values) == set(w)
        assert all(np.array_equal(store.values[k], w[k]) for k in w)
    wout = P.make_weights_out(4, 9, upRes=8, nets=(1, 2))
    cfg = on.make_cfg_out(4, upRes=8)
    for idx, firstGen in ((1, True), (2, False)):
        store = og.VarStore(seed=9)
        ctx = og.Context(store, torch.float32)
        spec = P.SHIPPED_8X[idx]
        with ctx.variable_scope("gen_%d" % idx):
            if firstGen:
                xin = torch.zeros(1, 4 * 4 * 6)
            else:
                xin = on.sampler_input_2(torch.zeros(1, 4 * 4 * 4), torch.zeros(1, 32 * 32), cfg)
            on.growing_gen(xin, ctx, cfg, currentUpres=3, output=True, firstGen=firstGen, filterSize=spec.filterSize,
                           startFms=spec.startFms, maxFms=spec.maxFms, add_adj_idcs=spec.add_adj_idcs,
                           first_nn_arch=spec.first_nn_arch, use_res_net=spec.use_res_net)
        assert set(store.values) == set(wout[idx])
        assert all(np.array_equal(store.values[k], wout[idx][k]) for k in wout[idx])


def test_layer_api_side_effects_and_quirks():
    G.reset_default_graph()
    x = G.placeholder([None, 8 * 8 * 3], "x")
    img = G.reshape(x, [-1, 8, 8, 3])
    gan = GAN(img)
    a, lin = gan.convolutional_layer(16, [3, 3], relu, name="c1")
    assert a.shape == (None, 8, 8, 16) and lin.node.op == "conv" and a.node.op == "act"
    assert gan.layer is a and gan.layer_num == 1 and gan.getDOFs() == 3 * 3 * 3 * 16 + 16
    # stride 2, k=4 (discriminator): ceil(H/2)
    d, _ = gan.convolutional_layer(8, [4, 4], lrelu, stride=[2], name="c2")
    assert d.shape == (None, 4, 4, 8)
    # depools ignore their argument and act on the cursor (App. D.1)
    other = G.reshape(G.placeholder([None, 2 * 2 * 1], "o"), [-1, 2, 2, 1])
    up = gan.max_depool(in_layer=other, height_factor=2, width_factor=2)
    assert up.shape == (None, 8, 8, 8) and up.node.inputs[0] is d
    # residual_block: B reads the cursor (= A), shortcut reads in_layer; returns (act, lin)
    r, rl = gan.residual_block(8, 4, [3, 3], relu, name="RB", in_layer=up)
    assert rl.node.op == "add" and r.node.op == "act"
    names = list(G.get_default_graph().variables)
    assert "RB_A/weight" in names and "RB_s/bias" in names
    # wscale constant = float32(gain / sqrt(fan_in)) (tools_wscale/GAN.py:664-668)
    wv = gan.weight_variable([5, 5, 4, 9], gain=1)
    assert wv["wscale"] == np.float32(1.0 / np.sqrt(100))
    with pytest.raises(ValueError):
        gan.convolutional_layer(4, [3, 3], activation_function=np.sin, name="bad")  # unknown activation callable


def test_unsupported_graph_fails_loudly():
    G.reset_default_graph()
    x = G.reshape(G.placeholder([None, 4 * 4 * 8], "x"), [-1, 4, 4, 8])
    gan = GAN(x)
    a, _ = gan.convolutional_layer(8, [3, 3], relu, name="c")
    y = G.relu(G.add(a, x))  # add of non-conv operands with 8 channels: no kernel for it
    w = W.init_graph_variables(G.get_default_graph(), 1)
    with pytest.raises(NotImplementedError):
        engine.CompiledNet(y, w, 2, dry=True)


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mpg.h")).read()
    declared = set(re.findall(r"\b(mpg_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mpg_handle", "mpg_conv_plan"}
    assert {"mpg_create", "mpg_conv_plan_create", "mpg_slice_assemble", "mpg_transpose3d", "mpg_pack_channels",
            "mpg_dens_residual", "mpg_threshold"} <= declared
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libmpg_b200.so does not export %s" % name
    assert lib.mpg_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MpgError):
        capi.Handle(0)


def test_pick_batch_and_slab_range():
    from mpgan_b200 import parallel as par
    assert P._pick_batch(512, 8) == 8 and P._pick_batch(20, 8) == 5 and P._pick_batch(3, 8) == 3
    assert par.slab_range(0, 4, 512) == (0, 128) and par.slab_range(3, 4, 512) == (384, 512)
    with pytest.raises(ValueError):
        par.slab_range(0, 3, 512)


def test_auto_batches_and_exchange_selection():
    """8x pipeline: pixels per launch stay ~constant across slice sizes; the fused peer-store exchange is only selected
    on CUDA with an initialised NCCL group (CPU / gloo runs use the all_to_all formulation)."""
    from mpgan_b200 import parallel as par
    assert P.auto_batches(512) == (32, 16, 16) and P.auto_batches(2048) == (2, 1, 1) and P.auto_batches(128) == (32, 16, 16)
    assert P.auto_batches(1024) == (8, 4, 4)
    assert par.p2p_usable(512, 1) is False
    if not torch.cuda.is_available():
        assert par.p2p_usable(512, 8) is False
        assert par.make_peer_slab((64, 512, 512), torch.device("cpu"), 512, 8) is None


def test_dry_run_predicts_the_tiny_kernel_for_the_tail_layers():
    """engine._predict_kind mirrors csrc/conv_plan.cu: ru3 of gen_resnet (8->2, 2(+8)->1) goes to the CUDA-core
    kernel, the 128->32 layer to the tap-folded tcgen05 kernel, the wide layers to the plain implicit GEMM."""
    G.reset_default_graph()
    cfg = N.config_4x(32, upsampling_mode=2)
    out = N.gen_resnet(G.placeholder([None, 32 * 32 * 4], "x"), cfg)
    w = W.init_graph_variables(G.get_default_graph(), 1)
    lab = _labels(engine.CompiledNet(out, w, 8, dry=True))
    kinds = {l.split()[1]: l.split()[0] for l in lab if l.startswith("conv")}
    assert kinds["g_cA3"] == "conv[ct]" and kinds["g_cB3+g_s3"] == "conv[ct]"
    assert kinds["g_cA2"] == "conv[nf]" and kinds["g_cB1+g_s1"] == "conv[tc]" and kinds["g_cA1"] == "conv[tc]"
