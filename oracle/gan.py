"""ORACLE (test infrastructure only - never imported by the product path).

Eager torch-CPU restatement of the reference layer library `class GAN` (tools_wscale/GAN.py) with
the same method names, arguments, return tuples and side effects (`self.layer` cursor, `DOFs`,
`layer_num`), including the quirks of SURVEY App. D.  TensorFlow variable scopes are emulated by a
scope stack; variables live in a `VarStore` keyed by the reference's variable names
(e.g. "generator/g_cA0/weight").  PARITY UNPINNED (no reference tests exist).
"""
import contextlib
import zlib

import numpy as np
import torch

from . import tf_ops


# ------------------------------------------------------------------ activations (callables with __name__)
def relu(x):
    return torch.relu(x)


def lrelu(x, leak=0.2, name="lrelu"):
    return tf_ops.lrelu(x, leak)


def tanh(x):
    return torch.tanh(x)


# ------------------------------------------------------------------ variable store / scopes
def init_variable(seed, name, shape, kind):
    """Deterministic initial value of variable `name` (independent of creation order).
    kind: 'normal' -> N(0,1) fp32 (tf.initializers.random_normal, tools_wscale/GAN.py:668);
    ('const', v) -> constant (bias 0.1 :683; BN gamma 1 / beta 0 / mean 0 / var 1)."""
    if kind == "normal":
        rng = np.random.default_rng(np.random.SeedSequence([int(seed), zlib.crc32(name.encode("utf-8"))]))
        return rng.standard_normal(size=tuple(shape), dtype=np.float32)
    assert kind[0] == "const"
    return np.full(tuple(shape), kind[1], dtype=np.float32)


class VarStore:
    def __init__(self, seed=1, values=None):
        self.seed = seed
        self.values = dict(values) if values else {}
        self.order = []

    def get(self, name, shape, kind):
        if name not in self.values:
            self.values[name] = init_variable(self.seed, name, shape, kind)
        v = self.values[name]
        assert tuple(v.shape) == tuple(shape), (name, v.shape, shape)
        if name not in self.order:
            self.order.append(name)
        return v


class Context:
    """Graph-level state TF keeps globally: variable store, scope stack, compute dtype, train flag."""

    def __init__(self, store, dtype=torch.float32):
        self.store = store
        self.dtype = dtype
        self.scopes = []
        self.bn_updates = {}  # name -> new moving stat (training only)

    @contextlib.contextmanager
    def variable_scope(self, name):
        self.scopes.append(name)
        try:
            yield
        finally:
            self.scopes.pop()

    def var(self, leaf, shape, kind):
        name = "/".join(self.scopes + [leaf])
        return torch.as_tensor(self.store.get(name, shape, kind)).to(self.dtype), name


class GAN(object):
    """tools_wscale/GAN.py:17-35."""

    def __init__(self, _image, ctx, bn_decay=0.999):
        self.layer = _image
        self.ctx = ctx
        self.batch_size = _image.shape[0]
        self.DOFs = 0
        self.preFlatShapes = []
        self.weight_stack = []
        self.layer_num = 0
        self.bn_decay = bn_decay

    # tools_wscale/GAN.py:80-119
    def convolutional_layer(self, outChannels, _patchShape, activation_function=tanh, stride=[1], name="conv",
                            reuse=False, batch_norm=False, train=None, in_layer=None, in_channels=None,
                            gain=np.sqrt(2)):
        if in_layer is None:
            in_layer = self.layer
        with self.ctx.variable_scope(name):
            self.layer_num += 1
            inChannels = int(in_channels) if in_channels is not None else int(in_layer.shape[-1])
            assert len(_patchShape) == 2, "only 2-D convolutions are on the hot path"
            W = self.weight_variable([_patchShape[0], _patchShape[1], inChannels, outChannels], name=name, gain=gain)
            self.layer = self.conv2d(in_layer, W, stride)
            self.DOFs += _patchShape[0] * _patchShape[1] * inChannels * outChannels
            self.weight_stack.append(W)
            b = self.bias_variable([outChannels], name=name)
            self.layer = self.layer + b
            self.DOFs += outChannels
            if batch_norm:
                self.layer = self._batch_norm(self.layer, train)
            layer_lin = self.layer
            if activation_function:
                self.layer = activation_function(self.layer)
            return self.layer, layer_lin

    def _batch_norm(self, x, train):
        """tf.contrib.layers.batch_norm in the conv's own scope (tools_wscale/GAN.py:110)."""
        c = x.shape[-1]
        beta, _ = self.ctx.var("beta", [c], ("const", 0.0))
        gamma, _ = self.ctx.var("gamma", [c], ("const", 1.0))
        mm, mm_name = self.ctx.var("moving_mean", [c], ("const", 0.0))
        mv, mv_name = self.ctx.var("moving_variance", [c], ("const", 1.0))
        if train:
            y, mean, var = tf_ops.batch_norm_training(x, gamma, beta)
            d = self.bn_decay
            # assign_moving_average: a scope that is evaluated twice in one step (the discriminator on real and on
            # generated input, GAN/multipassGAN-4x.py:729-730) updates the same variables twice; TF leaves the order
            # of the two UPDATE_OPS undefined - this oracle applies them in graph-construction order (real, fake)
            mm_cur = self.ctx.bn_updates.get(mm_name, mm.detach())
            mv_cur = self.ctx.bn_updates.get(mv_name, mv.detach())
            self.ctx.bn_updates[mm_name] = mm_cur * d + mean.detach() * (1 - d)
            self.ctx.bn_updates[mv_name] = mv_cur * d + var.detach() * (1 - d)
            return y
        return tf_ops.batch_norm_inference(x, gamma, beta, mm, mv)

    # tools_wscale/GAN.py:126-147
    def residual_block(self, s1, s2, filter, activation_function=tanh, name="RB", reuse=False, batch_norm=False,
                       train=None, in_layer=None):
        if in_layer is None:
            in_layer = self.layer
        filter1 = [1, 1]
        A, _ = self.convolutional_layer(s1, filter, activation_function, stride=[1], name=name + "_A",
                                        in_layer=in_layer, reuse=reuse, batch_norm=batch_norm, train=train)
        B, _ = self.convolutional_layer(s2, filter, None, stride=[1], name=name + "_B", reuse=reuse,
                                        batch_norm=batch_norm, train=train)
        s, _ = self.convolutional_layer(s2, filter1, None, stride=[1], name=name + "_s", in_layer=in_layer,
                                        reuse=reuse, batch_norm=batch_norm, train=train)
        self.layer = B + s
        layer_lin = self.layer
        if activation_function:
            self.layer = activation_function(self.layer)
        return self.layer, layer_lin

    # tools_wscale/GAN.py:423-435
    def flatten(self):
        s = self.layer.shape
        self.preFlatShapes.append(s)
        flatSize = int(s[1]) * int(s[2]) * int(s[3])
        self.layer = self.layer.reshape(-1, flatSize)
        return flatSize

    # tools_wscale/GAN.py:438-456
    def fully_connected_layer(self, _numHidden, _act, name="full", gain=np.sqrt(2)):
        with self.ctx.variable_scope(name):
            self.layer_num += 1
            numInput = int(self.layer.shape[1])
            W = self.weight_variable([numInput, _numHidden], name=name, gain=gain)
            b = self.bias_variable([_numHidden], name=name)
            self.DOFs += numInput * _numHidden + _numHidden
            self.layer = self.layer @ W + b
            if _act:
                self.layer = _act(self.layer)
            return self.layer

    # tools_wscale/GAN.py:472-474
    def pixel_norm(self, in_layer, epsilon=1e-8):
        self.layer = tf_ops.pixel_norm(in_layer, epsilon)
        return self.layer

    # tools_wscale/GAN.py:501-523 -- NOTE: acts on self.layer, the in_layer argument is ignored (App. D.1)
    def max_depool(self, in_layer=None, depth_factor=2, height_factor=2, width_factor=2):
        x = self.layer
        self.layer = tf_ops.resize_nearest(x, x.shape[1] * height_factor, x.shape[2] * width_factor)
        return self.layer

    # tools_wscale/GAN.py:528-552 -- acts on self.layer; mode 0 bilinear, 1 nearest, 2 bicubic
    def avg_depool(self, window_size=[1, 1], window_stride=[2, 2], mode=0, scale=[2]):
        x = self.layer
        if len(scale) == 1:
            oh, ow = x.shape[1] * scale[0], x.shape[2] * scale[0]
        else:
            oh, ow = x.shape[1] * scale[0], x.shape[2] * scale[1]
        if mode == 1:
            self.layer = tf_ops.resize_nearest(x, oh, ow)
        elif mode == 2:
            self.layer = tf_ops.resize_bicubic_tf1(x, oh, ow)
        else:
            self.layer = tf_ops.resize_bilinear_tf1(x, oh, ow)
        return self.layer

    def y(self):
        return self.layer

    def getDOFs(self):
        return self.DOFs

    # tools_wscale/GAN.py:661-678 -- runtime weight scaling: v * float32(gain / sqrt(fan_in))
    def weight_variable(self, shape, name="w", gain=np.sqrt(2), use_he=False, in_lay=None, use_wscale=True):
        if in_lay is None:
            in_lay = np.prod(shape[:-1])
        std = gain / np.sqrt(in_lay)
        v, _ = self.ctx.var("weight", shape, "normal")
        return v * torch.tensor(np.float32(std)).to(v.dtype)

    # tools_wscale/GAN.py:682-683
    def bias_variable(self, shape, name="b"):
        v, _ = self.ctx.var("bias", shape, ("const", 0.1))
        return v

    # tools_wscale/GAN.py:686-691
    def conv2d(self, x, W, stride=[1]):
        assert len(stride) == 1 or stride[0] == stride[1]
        return tf_ops.conv2d_same(x, W, stride[0])
