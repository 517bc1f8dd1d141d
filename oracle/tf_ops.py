"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the TensorFlow-1.x / scipy op semantics the reference's generator hot path relies
on.  TensorFlow and Keras are not installable in this image, so these functions
restate the *published* op definitions and are anchored on the reference call sites cited per
function.  OP ARITHMETIC UNPINNED against TensorFlow itself (not installable; the reference ships no vectors):
the pins are the known-answer tests in tests/test_oracle_ops.py and the independent numpy restatement in
tests/golden/tf1_numpy_shim.py that the reference's own layer code was executed on (tests/test_golden.py).

All tensors are NHWC torch CPU tensors; `dtype` float32 restates what TF fp32 computes, float64 is
the ground truth both the fp32 oracle and the CUDA path are measured against.
"""
import numpy as np
import scipy.ndimage
import torch
import torch.nn.functional as F


def same_padding(size, k, s):
    """tf.nn.conv2d padding="SAME" (tools_wscale/GAN.py:691): out=ceil(in/s),
    pad_total=max((out-1)*s+k-in,0), pad_before=floor(pad_total/2)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv2d_same(x, w, stride=1):
    """tf.nn.conv2d(x, W, [1,s,s,1], "SAME") - cross-correlation, NHWC x HWIO
    (tools_wscale/GAN.py:686-691)."""
    kh, kw = w.shape[0], w.shape[1]
    pt, pb = same_padding(x.shape[1], kh, stride)
    pl, pr = same_padding(x.shape[2], kw, stride)
    xi = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xi, w.permute(3, 2, 0, 1).contiguous(), stride=stride)
    return y.permute(0, 2, 3, 1).contiguous()


def batch_norm_inference(x, gamma, beta, moving_mean, moving_var, eps=1e-3):
    """tf.contrib.layers.batch_norm(..., scale=True, center=True, epsilon=0.001, fused=False,
    is_training=False) (tools_wscale/GAN.py:110): gamma*(x-mean)/sqrt(var+eps)+beta."""
    inv = gamma / torch.sqrt(moving_var + eps)
    return x * inv + (beta - moving_mean * inv)


def batch_norm_training(x, gamma, beta, eps=1e-3):
    """is_training=True branch: biased batch statistics over N,H,W. Returns (y, mean, var)."""
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    y = (x - mean) / torch.sqrt(var + eps) * gamma + beta
    return y, mean, var


def lrelu(x, leak=0.2):
    """tools_wscale/GAN.py:733-737: 0.5*(1+leak)*x + 0.5*(1-leak)*abs(x)."""
    f1 = 0.5 * (1 + leak)
    f2 = 0.5 * (1 - leak)
    return f1 * x + f2 * x.abs()


def pixel_norm(x, epsilon=1e-8):
    """tools_wscale/GAN.py:472-474: x * rsqrt(mean(x^2, axis=3, keepdims) + eps)."""
    return x * torch.rsqrt((x * x).mean(dim=3, keepdim=True) + epsilon)


def resize_nearest(x, out_h, out_w):
    """tf.image.resize_images(..., method=1) / keras.backend.resize_images with
    align_corners=False (tools_wscale/GAN.py:517,541; GAN/multipassGAN-out.py:357):
    src = floor(dst * in / out)."""
    in_h, in_w = x.shape[1], x.shape[2]
    iy = torch.floor(torch.arange(out_h, dtype=torch.float64) * (in_h / out_h)).long().clamp_(max=in_h - 1)
    ix = torch.floor(torch.arange(out_w, dtype=torch.float64) * (in_w / out_w)).long().clamp_(max=in_w - 1)
    return x[:, iy][:, :, ix]


_BICUBIC_TABLE = None


def _tf1_bicubic_table():
    """TF1 ResizeBicubic coefficient table: 1024 entries, Keys cubic with A=-0.75, rows
    [w(x), w(x+1)] with x = i/1024 (InitCoeffsTable in resize_bicubic_op.cc)."""
    global _BICUBIC_TABLE
    if _BICUBIC_TABLE is None:
        a = -0.75
        tab = np.zeros((1025, 2), dtype=np.float32)
        for i in range(1025):
            x = np.float32(i) / np.float32(1024)
            tab[i, 0] = ((a + 2) * x - (a + 3)) * x * x + 1
            x = x + np.float32(1.0)
            tab[i, 1] = ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
        _BICUBIC_TABLE = tab
    return _BICUBIC_TABLE


def _bicubic_axis_weights(in_size, out_size):
    """Per output index: 4 clamped source indices and 4 weights (legacy, no half-pixel centres,
    align_corners=False, no border renormalisation)."""
    tab = _tf1_bicubic_table()
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.zeros((out_size, 4), dtype=np.int64)
    wts = np.zeros((out_size, 4), dtype=np.float32)
    for o in range(out_size):
        in_f = np.float32(o) * scale
        i = int(np.floor(in_f))
        delta = in_f - np.float32(i)
        off = int(np.rint(delta * np.float32(1024)))  # lrintf(delta * kTableSize)
        wts[o, 0] = tab[off, 1]
        wts[o, 1] = tab[off, 0]
        wts[o, 2] = tab[1024 - off, 0]
        wts[o, 3] = tab[1024 - off, 1]
        for t in range(4):
            idx[o, t] = min(max(i - 1 + t, 0), in_size - 1)
    return idx, wts


def resize_bicubic_tf1(x, out_h, out_w):
    """tf.image.resize_images(..., method=2) of TF 1.x (tools_wscale/GAN.py:541 with mode=2 from
    GAN/multipassGAN-out.py:330).  SURVEY App. B.6: highest-risk oracle assumption (cannot be
    cross-checked against TF here)."""
    dt = x.dtype
    iy, wy = _bicubic_axis_weights(x.shape[1], out_h)
    ix, wx = _bicubic_axis_weights(x.shape[2], out_w)
    wy_t = torch.as_tensor(wy, dtype=dt)
    wx_t = torch.as_tensor(wx, dtype=dt)
    # TF interpolates along x first for each of the 4 rows, then along y
    rows = x[:, torch.as_tensor(iy)]  # [N, out_h, 4, W, C]
    cols = rows[:, :, :, torch.as_tensor(ix)]  # [N, out_h, 4, out_w, 4, C]
    along_x = (cols * wx_t.view(1, 1, 1, out_w, 4, 1)).sum(dim=4)  # [N, out_h, 4, out_w, C]
    return (along_x * wy_t.view(1, out_h, 4, 1, 1)).sum(dim=2)


def resize_bilinear_tf1(x, out_h, out_w):
    """tf.image.resize_images(..., method=0), align_corners=False, legacy (no half-pixel)."""
    dt = x.dtype

    def axis(in_size, out_size):
        scale = in_size / out_size
        f = np.arange(out_size, dtype=np.float64) * scale
        lo = np.floor(f).astype(np.int64)
        hi = np.minimum(lo + 1, in_size - 1)
        return torch.as_tensor(lo), torch.as_tensor(hi), torch.as_tensor(f - lo, dtype=dt)

    ylo, yhi, yl = axis(x.shape[1], out_h)
    xlo, xhi, xl = axis(x.shape[2], out_w)
    top = x[:, ylo]
    bot = x[:, yhi]
    rows = top + (bot - top) * yl.view(1, -1, 1, 1)
    left = rows[:, :, xlo]
    right = rows[:, :, xhi]
    return left + (right - left) * xl.view(1, 1, -1, 1)


def zoom_linear(a, zoom):
    """scipy.ndimage.zoom(a, zoom, order=1, mode='constant', cval=0.0) - the reference calls scipy
    itself (GAN/multipassGAN-out.py:401-421, GAN/multipassGAN-4x.py:1095-1103); so does the oracle."""
    return scipy.ndimage.zoom(np.asarray(a), zoom, order=1, mode="constant", cval=0.0)


def zoom_linear_axis_ref(a, axis, factor):
    """Explicit align-corners lerp along one axis (what zoom(order=1) computes);
    used to pin zoom_linear and the CUDA slice assembler."""
    a = np.asarray(a, dtype=np.float64)
    n_in = a.shape[axis]
    n_out = int(round(n_in * factor))
    if n_out == 1 or n_in == 1:
        coord = np.zeros(n_out)
    else:
        coord = np.arange(n_out, dtype=np.float64) * (n_in - 1) / (n_out - 1)
    lo = np.clip(np.floor(coord).astype(np.int64), 0, n_in - 1)
    hi = np.minimum(lo + 1, n_in - 1)
    t = coord - lo
    shape = [1] * a.ndim
    shape[axis] = n_out
    t = t.reshape(shape)
    return np.take(a, lo, axis=axis) * (1 - t) + np.take(a, hi, axis=axis) * t
