"""Structural pins of the oracle networks against the numbers derived from the reference source in
SURVEY App. A (parameter counts, FLOPs, variable names) and against reference quirks (App. D)."""
import numpy as np
import torch

from oracle import gan as og
from oracle import networks as on
from oracle import pipeline as op


def _count(store, pred):
    return sum(v.size for k, v in store.values.items() if pred(k))


def test_gen_resnet_dofs_and_names():
    store = og.VarStore(seed=1)
    ctx = og.Context(store, torch.float32)
    y, gan = on.gen_resnet(torch.zeros(1, 8 * 8 * 4), ctx, on.make_cfg_4x(8))
    assert y.shape == (1, 32 * 32)
    assert gan.getDOFs() == 634214  # 4x gen params
    assert _count(store, lambda k: k.endswith(("weight", "bias"))) == 634214
    # BN on ru1-ru3 (9 convs), not on ru4: 4 vectors per conv -> 1008 floats ("+1 008 BN")
    assert _count(store, lambda k: k.rsplit("/", 1)[1] in ("beta", "gamma", "moving_mean", "moving_variance")) == 2016
    assert "generator/g_cA0/weight" in store.values and "generator/g_s3/bias" in store.values
    assert "generator/g_cB2/moving_variance" in store.values and "generator/g_cA3/gamma" not in store.values
    assert store.values["generator/g_cB1/weight"].shape == (5, 5, 128, 128)
    assert float(y.min()) >= 0.0  # final relu


def test_growing_gen_dofs():
    cfg = on.make_cfg_out(8, upRes=8)
    store = og.VarStore(seed=1)
    ctx = og.Context(store, torch.float32)
    y, gan = on.growing_gen(torch.zeros(1, 8 * 8 * 6), ctx, cfg, currentUpres=3, output=True, firstGen=True, filterSize=3,
                            startFms=256, maxFms=256, add_adj_idcs=True, first_nn_arch=True, use_res_net=True)
    assert y.shape == (1, 64 * 64)
    assert _count(store, lambda k: True) == 1864961  # 8x net1
    assert "generator/genBlock2/g_cA_first/weight" in store.values and "generator/genBlock8/g_cdensOut8/weight" in store.values
    store2 = og.VarStore(seed=1)
    ctx2 = og.Context(store2, torch.float32)
    xin = on.sampler_input_2(torch.zeros(1, 8 * 8 * 4), torch.zeros(1, 64 * 64), cfg)
    on.growing_gen(xin, ctx2, cfg, currentUpres=3, output=True, firstGen=False, filterSize=5, startFms=192, maxFms=192,
                   first_nn_arch=False, use_res_net=True)
    assert _count(store2, lambda k: True) == 774301  # net2
    assert store2.values["generator/g_cA_1/weight"].shape == (5, 5, 5, 16)


def test_disc_binclass_dofs_and_shapes():
    cfg = on.make_cfg_4x(16)
    store = og.VarStore(seed=1)
    ctx = og.Context(store, torch.float32)
    logits, d1, d2, d3, d4 = on.disc_binclass(torch.zeros(2, 16 * 16 * 4), torch.zeros(2, 64 * 64), ctx, cfg)
    assert logits.shape == (2, 1) and d1.shape == (2, 32, 32, 32) and d2.shape == (2, 16, 16, 64)
    assert d3.shape == (2, 8, 8, 128) and d4.shape == (2, 8, 8, 256)
    assert _count(store, lambda k: k.endswith(("weight", "bias"))) == 706017


def test_depool_ignores_its_argument_quirk():
    """App. D.1: max_depool / avg_depool act on self.layer, not on in_layer."""
    ctx = og.Context(og.VarStore(seed=1), torch.float32)
    g = og.GAN(torch.ones(1, 2, 2, 1), ctx)
    other = torch.full((1, 2, 2, 1), 7.0)
    out = g.max_depool(in_layer=other, height_factor=2, width_factor=2)
    assert out.shape == (1, 4, 4, 1) and float(out.max()) == 1.0


def test_pipeline_identity_networks_axis_bookkeeping():
    """With networks that return (a nearest-upsampled copy of) one input channel the whole multi-pass
    pipeline must reduce to known volumes: pass 1 alone on channel 0 == nearest xy-upsample of the
    z-lerped density, in canonical [Z,Y,X] order."""
    L, u = 4, 2
    S = L * u
    rng = np.random.default_rng(0)
    x = rng.random((L, L, L, 4)).astype(np.float32)

    def net1(rows):  # rows: [B, L*L*4] -> [B, S*S] nearest x u of channel 0
        b = rows.reshape(-1, L, L, 4)[..., 0]
        return b.repeat(u, axis=1).repeat(u, axis=2).reshape(-1, S * S)

    out = op.out_generate3d(x, u, net1, None, None, transposeAxis=0, threshold=False)
    import scipy.ndimage
    exp = scipy.ndimage.zoom(x[..., 0], [u, 1, 1], order=1).repeat(u, axis=1).repeat(u, axis=2)
    assert np.abs(out - exp).max() < 1e-6

    # pass 2 identity on y: output == pass-1 volume (the transposes must cancel)
    def net2(rows, yrows):
        return yrows

    out2 = op.out_generate3d(x, u, net1, net2, None, transposeAxis=0, threshold=False)
    assert np.abs(out2 - exp).max() < 1e-6
    out3 = op.out_generate3d(x, u, net1, net2, net2, transposeAxis=0, threshold=False)
    assert np.abs(out3 - exp).max() < 1e-6

    # pass 2 returning its (nearest-upsampled) velocity channel 1 reveals the vx<->vz swap (App. C):
    def net2v(rows, yrows):
        b = rows.reshape(-1, L, L, 4)[..., 1]
        return b.repeat(u, axis=1).repeat(u, axis=2).reshape(-1, S * S)

    out4 = op.out_generate3d(x, u, net1, net2v, None, transposeAxis=0, threshold=False)
    # slices along x of (y,z) planes of vz (channel 3), x-lerped; back in [Z,Y,X]
    vz = scipy.ndimage.zoom(x[..., 3], [1, 1, u], order=1)  # [Z,Y,Xu]
    exp4 = vz.repeat(u, axis=0).repeat(u, axis=1)
    assert np.abs(out4 - exp4).max() < 1e-6


def test_two_pass_4x_identity_networks():
    L, u = 2, 4
    S = L * u
    x = np.random.default_rng(1).random((L, L, L, 4)).astype(np.float32) + 0.1

    def net1(rows):
        b = rows.reshape(-1, L, L, 4)[..., 0]
        return b.repeat(u, axis=1).repeat(u, axis=2).reshape(-1, S * S)

    def net2(rows):  # returns the first-pass density channel
        return rows.reshape(-1, S, S, 4)[..., 0].reshape(-1, S * S)

    out, p1 = op.two_pass_4x(net1, net2, x, u, return_intermediate=True)
    assert np.array_equal(out, p1)  # transposes of pass 2 cancel exactly

    def net2vx(rows):  # channel 3 after the swaps is vx * u (App. C, 4x pass 2: (d, vy, vz, vx))
        return rows.reshape(-1, S, S, 4)[..., 3].reshape(-1, S * S)

    out = op.two_pass_4x(net1, net2vx, x, u)
    import scipy.ndimage
    exp = scipy.ndimage.zoom(x[..., 1] * u, [u, u, u], order=1)
    exp[exp < 0.0005] = 0
    assert np.abs(out - exp).max() < 1e-5


def test_tiled_apply_needs_exactly_the_receptive_field_halo():
    """SURVEY App. A.6 on the oracle (fp64): gen_resnet applied to overlapping tiles and stitched with
    `stitch_overlap` equals the whole-slice apply when the halo is the receptive-field radius (16 output pixels = 4
    low-res pixels: 8 convs of k=5), and differs when it is one low-res pixel short. This is the property the GPU
    pipeline's tiled mode (pipeline._TiledPassNet) relies on."""
    from oracle import tiles as ot_
    from oracle_nets import oracle_gen_resnet
    import mpgan_b200  # noqa: F401
    from mpgan_b200 import pipeline as P
    L, u = 24, 4
    w1, _ = P.make_weights_4x(L, 3, upRes=u, randomize_bn=True)
    rng = np.random.default_rng(5)
    x = rng.random((2, L, L, 4))
    whole = oracle_gen_resnet(w1, L, 2)(x.reshape(2, -1)).reshape(2, L * u, L * u, 1)
    for halo, exact in ((4, True), (3, False)):
        core = 8 if halo == 4 else 6           # (L - 2*halo) is a multiple of core in both cases
        T = core + 2 * halo
        nt = (L - 2 * halo) // core
        tiles = ot_.cut_overlap(x, T, halo)
        assert tiles.shape[0] == 2 * nt * nt
        yt = oracle_gen_resnet(w1, T, 2)(tiles.reshape(tiles.shape[0], -1)).reshape(-1, T * u, T * u, 1)
        got = ot_.stitch_overlap(yt, 2, nt, nt, halo * u)
        err = float(np.abs(got - whole).max())
        if exact:
            assert err < 1e-10, err
        else:
            assert err > 1e-6, err
