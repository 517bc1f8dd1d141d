"""Eager numpy stand-in for the slice of the TensorFlow-1.x / Keras-backend API that the reference's
generator code touches (tools_wscale/GAN.py, the net builders of GAN/multipassGAN-out.py and
GAN/multipassGAN-4x.py).  TEST INFRASTRUCTURE: used only by tests/golden/make_golden.py, which executes
the reference's OWN Python (read from /root/reference at generation time, never copied) on top of this
shim to produce golden vectors.  TensorFlow itself is not installable in this image, so:

  * what the fixtures PIN against the reference: layer wiring, variable names/shapes, wscale constants,
    bias/BN/activation order, cursor quirks, reshapes/concats/slices, channel and axis bookkeeping;
  * what stays a restatement of published TF semantics: the arithmetic of the ops below
    (conv2d SAME, contrib batch_norm inference, resize_images nearest / legacy bicubic).  They are written
    here independently of oracle/tf_ops.py (direct numpy tap loops in float64), so the two restatements
    check each other.

All tensors are float64 numpy arrays wrapped in `T`.
"""
import contextlib
import types

import numpy as np


class _Shape(list):
    def as_list(self):
        return list(self)


class T:
    """Eager tensor: numpy array + the handful of tf.Tensor methods the reference calls."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    def get_shape(self):
        return _Shape(self.a.shape)

    @property
    def shape(self):
        return _Shape(self.a.shape)

    def __add__(self, o):
        return T(self.a + _v(o))

    __radd__ = __add__

    def __sub__(self, o):
        return T(self.a - _v(o))

    def __mul__(self, o):
        return T(self.a * _v(o))

    __rmul__ = __mul__

    def __abs__(self):
        return T(np.abs(self.a))


def _v(x):
    return x.a if isinstance(x, T) else x


# ----------------------------------------------------------------------------- variables / scopes
class _State:
    def __init__(self):
        self.scopes = []
        self.values = {}      # injected: full variable name -> array
        self.requested = {}   # name -> shape, in creation order


STATE = _State()


def reset(values):
    STATE.scopes = []
    STATE.values = dict(values)
    STATE.requested = {}


class _Scope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    if isinstance(name, _Scope):  # tf.variable_scope(tf.get_variable_scope()) re-enters the current scope
        yield name
        return
    STATE.scopes.append(name)
    try:
        yield _Scope("/".join(STATE.scopes))
    finally:
        STATE.scopes.pop()


def get_variable_scope():
    return _Scope("/".join(STATE.scopes))


def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True):
    full = "/".join(STATE.scopes + [name])
    shape = tuple(int(s) for s in shape)
    STATE.requested[full] = shape
    if full not in STATE.values:
        raise KeyError("reference requested variable '%s' %s that the injected weights do not hold" % (full, shape))
    val = np.asarray(STATE.values[full])
    if val.shape != shape:
        raise ValueError("variable %s: reference shape %s, injected %s" % (full, shape, val.shape))
    return T(val)


AUTO_REUSE = "AUTO_REUSE"
float32 = "float32"
int32 = "int32"
bool = "bool"  # noqa: A001  (tf.bool)


def constant(value, name=None, dtype=None):
    if isinstance(value, (list, tuple)):
        return list(value)
    return value  # python / numpy scalar (np.float32(std) keeps its fp32 rounding: tools_wscale/GAN.py:667)


def constant_initializer(value, dtype=None):
    return ("const", value)


initializers = types.SimpleNamespace(random_normal=lambda dtype=None: ("normal",))
keras = types.SimpleNamespace(initializers=types.SimpleNamespace(he_normal=lambda dtype=None: ("he",)))


def shape(x):
    return list(x.a.shape)


def cast(x, dtype):
    return x


def reshape(x, shape):
    return T(np.reshape(_v(x), [int(s) for s in shape]))


def concat(values, axis):
    return T(np.concatenate([_v(v) for v in values], axis=axis))


def slice(x, begin, size):  # noqa: A001
    a = _v(x)
    idx = tuple(np.s_[b:(a.shape[i] if s == -1 else b + s)] for i, (b, s) in enumerate(zip(begin, size)))
    return T(a[idx])


def add(a, b):
    return T(_v(a) + _v(b))


def square(x):
    return T(_v(x) ** 2)


def rsqrt(x):
    return T(1.0 / np.sqrt(_v(x)))


def reduce_mean(x, axis=None, keep_dims=False, keepdims=False):
    return T(np.mean(_v(x), axis=axis, keepdims=keep_dims or keepdims))


def matmul(a, b):
    return T(_v(a) @ _v(b))


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2, out


def _conv2d(x, w, strides, padding):
    """tf.nn.conv2d: NHWC x HWIO cross-correlation, SAME padding split floor/ceil."""
    assert padding == "SAME" and strides[0] == 1 and strides[3] == 1
    x, w = _v(x), _v(w)
    sh, sw = strides[1], strides[2]
    kh, kw, cin, cout = w.shape
    n, h, wd, c = x.shape
    assert c == cin, (x.shape, w.shape)
    pt, pb, oh = _same_pad(h, kh, sh)
    pl, pr, ow = _same_pad(wd, kw, sw)
    xp = np.zeros((n, h + pt + pb, wd + pl + pr, c))
    xp[:, pt:pt + h, pl:pl + wd] = x
    out = np.zeros((n, oh, ow, cout))
    for dy in range(kh):
        for dx in range(kw):
            win = xp[:, dy:dy + (oh - 1) * sh + 1:sh, dx:dx + (ow - 1) * sw + 1:sw, :]
            out += win @ w[dy, dx]
    return T(out)


def _relu(x):
    return T(np.maximum(_v(x), 0.0))


_relu.__name__ = "relu"


def _tanh(x):
    return T(np.tanh(_v(x)))


_tanh.__name__ = "tanh"


def _sigmoid(x):
    return T(1.0 / (1.0 + np.exp(-_v(x))))


nn = types.SimpleNamespace(conv2d=lambda x, W, strides, padding: _conv2d(x, W, strides, padding), relu=_relu,
                           tanh=_tanh, sigmoid=_sigmoid)


def _batch_norm(x, decay=0.999, center=True, scale=False, epsilon=0.001, scope=None, reuse=None, fused=None,
                is_training=True):
    """tf.contrib.layers.batch_norm, inference branch (moving statistics); variables live in `scope`."""
    if is_training not in (False,):
        raise NotImplementedError("shim: only is_training=False is evaluated (apply path)")
    c = int(_v(x).shape[-1])
    with variable_scope(scope if scope is not None else "BatchNorm"):
        beta = get_variable("beta", [c]) if center else T(np.zeros(c))
        gamma = get_variable("gamma", [c]) if scale else T(np.ones(c))
        mean = get_variable("moving_mean", [c])
        var = get_variable("moving_variance", [c])
    return T((_v(x) - mean.a) / np.sqrt(var.a + epsilon) * gamma.a + beta.a)


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(batch_norm=_batch_norm))


def _nearest(a, oh, ow):
    """ResizeNearestNeighbor, align_corners=False: src = floor(dst * in / out)."""
    n, h, w, c = a.shape
    iy = np.minimum((np.arange(oh) * (h / oh)).astype(np.int64), h - 1)
    ix = np.minimum((np.arange(ow) * (w / ow)).astype(np.int64), w - 1)
    return a[:, iy][:, :, ix]


def _cubic_w(t, A=-0.75):
    t = np.abs(t)
    return np.where(t <= 1, ((A + 2) * t - (A + 3)) * t * t + 1, ((A * t - 5 * A) * t + 8 * A) * t - 4 * A)


def _bicubic_axis(a, out, axis):
    """TF1 ResizeBicubic along one axis (align_corners=False, no half-pixel centres, Keys A=-0.75, weights
    from a 1024-entry table, clamped taps, no renormalisation)."""
    n_in = a.shape[axis]
    scale = n_in / out
    res = np.zeros(a.shape[:axis] + (out,) + a.shape[axis + 1:])
    a = np.moveaxis(a, axis, 0)
    r = np.moveaxis(res, axis, 0)
    for o in range(out):
        f = o * scale
        i = int(np.floor(f))
        d = np.floor((f - i) * 1024 + 0.5) / 1024.0  # table lookup: delta quantised to 1/1024
        ws = [_cubic_w(1 + d), _cubic_w(d), _cubic_w(1 - d), _cubic_w(2 - d)]
        for k, wgt in zip((-1, 0, 1, 2), ws):
            r[o] += np.float32(wgt).astype(np.float64) * a[min(max(i + k, 0), n_in - 1)]
    return np.moveaxis(r, 0, axis)


def _resize_images(x, size, method=0, align_corners=False):
    a = _v(x)
    oh, ow = int(_v(size)[0]), int(_v(size)[1])
    if method == 1:
        return T(_nearest(a, oh, ow))
    if method == 2:
        return T(_bicubic_axis(_bicubic_axis(a, ow, 2), oh, 1))
    raise NotImplementedError("shim: resize method %r" % (method,))


image = types.SimpleNamespace(resize_images=_resize_images)


# keras.backend.resize_images(x, h_factor, w_factor, 'channels_last'): nearest repeat (tools_wscale/GAN.py:517)
def kb_resize_images(x, height_factor, width_factor, data_format, interpolation="nearest"):
    assert data_format == "channels_last"
    return T(np.repeat(np.repeat(_v(x), height_factor, axis=1), width_factor, axis=2))


keras_backend = types.SimpleNamespace(resize_images=kb_resize_images)


# ---- ops used by the 8x trainer's growing_disc (GAN/multipassGAN-8x.py:596-597, 752-866; tools_wscale/GAN.py:162-169)
def zeros_like(x):
    return T(np.zeros_like(_v(x)))


def clip_by_value(x, lo, hi):
    return T(np.clip(_v(x), lo, hi))


def _avg_pool(x, ksize, strides, padding):
    assert padding == "VALID" and list(ksize) == list(strides) and ksize[0] == 1 and ksize[3] == 1
    a = _v(x)
    kh, kw = int(ksize[1]), int(ksize[2])
    n, h, w, c = a.shape
    a = a[:, :h // kh * kh, :w // kw * kw]
    return T(a.reshape(n, h // kh, kh, w // kw, kw, c).mean(axis=(2, 4)))


nn.avg_pool = _avg_pool


# ---- ops used by tensorResample (GAN/multipassGAN-8x.py:545-594): index arithmetic + tf.gather_nd
int32 = "int32"
float32 = "float32"


def _cmp(self, o):
    return T(self.a > _v(o))


def _and(self, o):
    return T(np.logical_and(self.a != 0, np.asarray(_v(o)) != 0))


T.__gt__ = _cmp
T.__and__ = _and
T.__rsub__ = lambda self, o: T(_v(o) - self.a)


@contextlib.contextmanager
def name_scope(name):
    yield name


def floor(x):
    return T(np.floor(_v(x)))


def ones_like(x, dtype=None):
    return T(np.ones_like(_v(x)))


def where(cond, a, b):
    return T(np.where(_v(cond) != 0, _v(a), _v(b)))


def abs(x):  # noqa: A001
    return T(np.abs(_v(x)))


def reduce_prod(x, axis=None, keep_dims=False, keepdims=False):
    return T(np.prod(_v(x), axis=axis, keepdims=keep_dims or keepdims))


def cumsum(x, axis=0, exclusive=False):
    a = _v(x)
    c = np.cumsum(a, axis=axis)
    return T(c - a if exclusive else c)


def gather_nd(params, indices):
    """tf.gather_nd with full-rank-minus-channels indices [..., k]: params[idx[..., 0], ..., idx[..., k-1]] (trailing params
    axes are kept). Out-of-range indices yield 0, which is what the TensorFlow GPU kernel does (the CPU kernel raises)."""
    p, idx = _v(params), np.asarray(_v(indices)).astype(np.int64)
    k = idx.shape[-1]
    ok = np.ones(idx.shape[:-1], bool)
    for d in range(k):
        ok &= (idx[..., d] >= 0) & (idx[..., d] < p.shape[d])
    safe = np.where(ok[..., None], idx, 0)
    out = p[tuple(safe[..., d] for d in range(k))]
    return T(out * ok.reshape(ok.shape + (1,) * (out.ndim - ok.ndim)))
