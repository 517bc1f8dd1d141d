"""Known-answer tests pinning the oracle's TF-1.x / scipy op restatements.
The reference ships no tests or golden vectors (parity unpinned): these KATs ARE the pins."""
import numpy as np
import scipy.ndimage
import torch

from oracle import gan as og
from oracle import tf_ops


def test_same_padding_rules():
    # odd k, stride 1: symmetric; k=4 s=2 even input: 1/1; k=4 s=1: 1 before, 2 after (disc d_c4)
    assert tf_ops.same_padding(64, 5, 1) == (2, 2)
    assert tf_ops.same_padding(64, 3, 1) == (1, 1)
    assert tf_ops.same_padding(64, 1, 1) == (0, 0)
    assert tf_ops.same_padding(64, 4, 2) == (1, 1)
    assert tf_ops.same_padding(8, 4, 1) == (1, 2)
    assert tf_ops.same_padding(7, 4, 2) == (1, 2)  # out=4, total=(3*2+4-7)=3


def test_conv2d_same_is_cross_correlation_hwio():
    x = torch.zeros(1, 5, 5, 1)
    x[0, 2, 2, 0] = 1.0
    w = torch.arange(9, dtype=torch.float32).reshape(3, 3, 1, 1)
    y = tf_ops.conv2d_same(x, w)[0, :, :, 0]
    # impulse response of a cross-correlation is the FLIPPED kernel
    assert torch.equal(y[1:4, 1:4], torch.flip(w[:, :, 0, 0], dims=(0, 1)))
    # asymmetric even-kernel padding: k=4, s=1 on a 4x4 of ones -> corner sums follow pad (1 before, 2 after)
    x = torch.ones(1, 4, 4, 1)
    w = torch.ones(4, 4, 1, 1)
    y = tf_ops.conv2d_same(x, w)[0, :, :, 0]
    assert y[0, 0] == 9 and y[3, 3] == 4 and y[1, 1] == 16 and y.shape == (4, 4)
    # stride 2, k=4
    y2 = tf_ops.conv2d_same(torch.ones(1, 8, 8, 1), w, stride=2)[0, :, :, 0]
    assert y2.shape == (4, 4) and y2[0, 0] == 9 and y2[1, 1] == 16 and y2[3, 3] == 9


def test_conv2d_channels_hwio():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((2, 6, 7, 3)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((3, 3, 3, 4)).astype(np.float32))
    y = tf_ops.conv2d_same(x, w)
    # direct loop for one output element
    n, i, j, co = 1, 2, 3, 2
    acc = 0.0
    for di in range(3):
        for dj in range(3):
            for ci in range(3):
                ii, jj = i + di - 1, j + dj - 1
                if 0 <= ii < 6 and 0 <= jj < 7:
                    acc += float(x[n, ii, jj, ci]) * float(w[di, dj, ci, co])
    assert abs(float(y[n, i, j, co]) - acc) < 1e-5


def test_batch_norm_inference_fresh_init_is_div_sqrt_1p001():
    x = torch.tensor([[[[1.0, -2.0]]]])
    y = tf_ops.batch_norm_inference(x, torch.ones(2), torch.zeros(2), torch.zeros(2), torch.ones(2))
    assert torch.allclose(y, x / np.sqrt(1.001))
    y = tf_ops.batch_norm_inference(x, torch.tensor([2.0, 3.0]), torch.tensor([0.5, -0.5]), torch.tensor([1.0, 1.0]),
                                    torch.tensor([3.0, 0.25]))
    exp = torch.tensor([[[[2 * (1 - 1) / np.sqrt(3.001) + 0.5, 3 * (-2 - 1) / np.sqrt(0.251) - 0.5]]]])
    assert torch.allclose(y, exp.float())


def test_lrelu_and_pixel_norm():
    x = torch.tensor([-2.0, 0.0, 3.0])
    assert torch.allclose(tf_ops.lrelu(x), torch.tensor([-0.4, 0.0, 3.0]))  # == max(x, 0.2x)
    assert torch.allclose(og.lrelu(x), torch.tensor([-0.4, 0.0, 3.0]))
    p = torch.tensor([[[[3.0, 4.0]]]])
    y = tf_ops.pixel_norm(p)
    assert torch.allclose(y, p / np.sqrt(12.5 + 1e-8))
    assert torch.allclose(tf_ops.pixel_norm(torch.zeros(1, 1, 1, 4)), torch.zeros(1, 1, 1, 4))


def test_resize_nearest_integer_factor_is_repetition():
    x = torch.arange(6, dtype=torch.float32).reshape(1, 2, 3, 1)
    y = tf_ops.resize_nearest(x, 4, 6)
    assert torch.equal(y[0, :, :, 0], x[0, :, :, 0].repeat_interleave(2, 0).repeat_interleave(2, 1))
    y8 = tf_ops.resize_nearest(x, 16, 24)
    assert torch.equal(y8[0, :, :, 0], x[0, :, :, 0].repeat_interleave(8, 0).repeat_interleave(8, 1))


def test_tf1_bicubic_known_answers():
    tab = tf_ops._tf1_bicubic_table()
    assert tab[0, 0] == 1.0 and tab[0, 1] == 0.0  # w(0)=1, w(1)=0
    assert abs(tab[512, 0] - 0.59375) < 1e-7 and abs(tab[512, 1] + 0.09375) < 1e-7  # Keys A=-0.75 at 0.5 / 1.5
    # constants are preserved in the interior (weights sum to 1) and on the borders (clamped taps)
    c = torch.full((1, 5, 5, 1), 3.0)
    assert torch.allclose(tf_ops.resize_bicubic_tf1(c, 20, 20), torch.full((1, 20, 20, 1), 3.0), atol=1e-6)
    # out index multiple of the factor hits the source sample exactly (delta = 0): legacy, no half-pixel shift
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.random((1, 6, 6, 1)).astype(np.float32))
    y = tf_ops.resize_bicubic_tf1(x, 24, 24)
    assert torch.allclose(y[0, ::4, ::4, 0], x[0, :, :, 0], atol=1e-6)
    # 1-D ramp, x2: halfway sample = -0.09375*f(i-1) + 0.59375*f(i) + 0.59375*f(i+1) - 0.09375*f(i+2)
    r = torch.arange(8, dtype=torch.float32).reshape(1, 1, 8, 1).repeat(1, 2, 1, 1)
    y = tf_ops.resize_bicubic_tf1(r, 2, 16)
    assert abs(float(y[0, 0, 5, 0]) - 2.5) < 1e-6
    assert abs(float(y[0, 0, 1, 0]) - (-0.09375 * 0 + 0.59375 * 0 + 0.59375 * 1 - 0.09375 * 2)) < 1e-6  # left clamp
    # differs from torch's half-pixel bicubic
    t = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(24, 24), mode="bicubic", align_corners=False)
    assert float((t.permute(0, 2, 3, 1) - tf_ops.resize_bicubic_tf1(x, 24, 24)).abs().max()) > 1e-3


def test_zoom_linear_is_align_corners_lerp():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((5, 4, 3, 2)).astype(np.float32)
    for axis, z in ((0, [4, 1, 1, 1]), (1, [1, 8, 1, 1]), (2, [1, 1, 4, 1])):
        got = tf_ops.zoom_linear(a, z)
        ref = tf_ops.zoom_linear_axis_ref(a, axis, z[axis])
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-6
    # default mode == mode='constant' for these inputs (reference calls both spellings)
    assert np.array_equal(scipy.ndimage.zoom(a, [4, 1, 1, 1], order=1),
                          scipy.ndimage.zoom(a, [4, 1, 1, 1], order=1, mode="constant", cval=0.0))


def test_wscale_and_bias_init():
    store = og.VarStore(seed=3)
    ctx = og.Context(store, torch.float32)
    g = og.GAN(torch.zeros(1, 4, 4, 3), ctx)
    with ctx.variable_scope("s"):
        w = g.weight_variable([5, 5, 3, 7], gain=np.sqrt(2))
        b = g.bias_variable([7])
    raw = store.values["s/weight"]
    assert raw.dtype == np.float32 and abs(raw.std() - 1.0) < 0.1
    assert np.array_equal(w.numpy(), raw * np.float32(np.sqrt(2) / np.sqrt(75)))
    assert np.all(b.numpy() == np.float32(0.1))
    with ctx.variable_scope("fc"):
        wf = g.weight_variable([16384, 1])  # FC: fan-in = numInput (tools_wscale/GAN.py:444)
    assert np.array_equal(wf.numpy(), store.values["fc/weight"] * np.float32(np.sqrt(2) / np.sqrt(16384)))
