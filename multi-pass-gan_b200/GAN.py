"""Drop-in for the reference layer library `tools_wscale/GAN.py` (class GAN + lrelu).

Same method names, argument lists, return tuples and side effects (the `self.layer` cursor with the
quirks of SURVEY App. D, `DOFs`, `layer_num`, `weight_stack`) as the reference; instead of emitting
TensorFlow ops each call records a node in the graph IR (graph.py) that engine.py lowers to fused
tcgen05 / CUDA-core kernels.  Activations are passed as callables and identified by `__name__`
(`relu`, `lrelu`, `tanh`) exactly as the reference prints them (tools_wscale/GAN.py:116).
"""
import numpy as np

from . import graph as G
from .graph import lrelu, relu, tanh  # noqa: F401  (re-exported: `from GAN import GAN, lrelu`)

_ACT_NAMES = {"relu": "relu", "lrelu": "lrelu", "tanh": "tanh"}


def _act_kind(fn):
    if fn is None:
        return None
    name = getattr(fn, "__name__", None)
    if name not in _ACT_NAMES:
        raise ValueError("unsupported activation_function %r (use relu, lrelu, tanh or None)" % (fn,))
    return _ACT_NAMES[name]


class GAN(object):
    # tools_wscale/GAN.py:19-35
    def __init__(self, _image, bn_decay=0.999):
        self.layer = _image
        self.batch_size = None  # symbolic (tf.shape(_image)[0] in the reference)
        self.DOFs = 0
        self.preFlatShapes = []
        self.weight_stack = []
        self.layer_num = 0
        self.layer_num_gen = 0
        self.layer_num_disc = 0
        self.bn_decay = bn_decay
        self.verbose = False
        self._g = _image.graph

    def _print(self, msg):
        if self.verbose:
            print(msg)

    # tools_wscale/GAN.py:80-119
    def convolutional_layer(self, outChannels, _patchShape, activation_function=tanh, stride=[1], name="conv",
                            reuse=False, batch_norm=False, train=None, in_layer=None, in_channels=None,
                            gain=np.sqrt(2)):
        if in_layer is None:
            in_layer = self.layer
        with G.variable_scope(name, reuse=reuse):
            self.layer_num += 1
            if in_channels is not None:
                inChannels = int(in_channels)
            else:
                inChannels = int(in_layer.get_shape()[-1])
            if len(_patchShape) != 2:
                raise NotImplementedError("3-D convolutions are outside the accelerated path (SURVEY §2 #1)")
            W = self.weight_variable([_patchShape[0], _patchShape[1], inChannels, outChannels], name=name, gain=gain)
            self.DOFs += _patchShape[0] * _patchShape[1] * inChannels * outChannels
            self.weight_stack.append(W)
            b = self.bias_variable([outChannels], name=name)
            self.DOFs += outChannels
            bn = None
            if batch_norm:
                # tf.contrib.layers.batch_norm(scope=current scope): beta/gamma/moving_* live next to weight/bias
                bn = dict(beta=self._g.get_variable("beta", [outChannels], ("const", 0.0)),
                          gamma=self._g.get_variable("gamma", [outChannels], ("const", 1.0)),
                          moving_mean=self._g.get_variable("moving_mean", [outChannels], ("const", 0.0)),
                          moving_variance=self._g.get_variable("moving_variance", [outChannels], ("const", 1.0)),
                          decay=self.bn_decay, train=train)
            self.layer = self.conv2d(in_layer, W, stride, _bias=b, _bn=bn)
            layer_lin = self.layer
            kind = _act_kind(activation_function)
            if kind:
                self.layer = G.activation(self.layer, kind)
            self._print("Convolutional Layer '{}' {} ({}) : {}, BN:{}".format(
                name, W["var"].shape, kind or "None", self.layer.get_shape(), batch_norm))
            return self.layer, layer_lin

    # tools_wscale/GAN.py:126-147
    def residual_block(self, s1, s2, filter, activation_function=tanh, name="RB", reuse=False, batch_norm=False,
                       train=None, in_layer=None):
        if in_layer is None:
            in_layer = self.layer
        if len(filter) == 2:
            filter1 = [1, 1]
        else:
            raise NotImplementedError("3-D residual blocks are outside the accelerated path")
        self._print("Residual Block:")
        A, _ = self.convolutional_layer(s1, filter, activation_function, stride=[1], name=name + "_A",
                                        in_layer=in_layer, reuse=reuse, batch_norm=batch_norm, train=train)
        B, _ = self.convolutional_layer(s2, filter, None, stride=[1], name=name + "_B", reuse=reuse,
                                        batch_norm=batch_norm, train=train)
        s, _ = self.convolutional_layer(s2, filter1, None, stride=[1], name=name + "_s", in_layer=in_layer,
                                        reuse=reuse, batch_norm=batch_norm, train=train)
        self.layer = G.add(B, s)
        layer_lin = self.layer
        kind = _act_kind(activation_function)
        if kind:
            self.layer = G.activation(self.layer, kind)
        return self.layer, layer_lin

    # tools_wscale/GAN.py:472-474
    def pixel_norm(self, in_layer, epsilon=1e-8):
        self.layer = G.pixel_norm(in_layer, epsilon)
        return self.layer

    # tools_wscale/GAN.py:501-523 -- acts on self.layer; `in_layer` is ignored by the reference (App. D.1)
    def max_depool(self, in_layer=None, depth_factor=2, height_factor=2, width_factor=2):
        s = self.layer.get_shape()
        if len(s) != 4:
            raise NotImplementedError("3-D depool is outside the accelerated path")
        self.layer = G.resize_images(self.layer, [s[1] * height_factor, s[2] * width_factor], 1)
        self._print("Max Depool : {}".format(self.layer.get_shape()))
        return self.layer

    # tools_wscale/GAN.py:528-552 -- acts on self.layer; mode 0 bilinear, 1 nearest, 2 bicubic
    def avg_depool(self, window_size=[1, 1], window_stride=[2, 2], mode=0, scale=[2]):
        s = self.layer.get_shape()
        if len(s) != 4:
            raise NotImplementedError("3-D depool is outside the accelerated path")
        if len(scale) == 1:
            outWidth, outHeight = s[2] * scale[0], s[1] * scale[0]
        else:
            outWidth, outHeight = s[2] * scale[1], s[1] * scale[0]
        self.layer = G.resize_images(self.layer, [int(outHeight), int(outWidth)], mode)
        self._print("Avg Depool {}: {}".format(window_size, self.layer.get_shape()))
        return self.layer

    # tools_wscale/GAN.py:423-435
    def flatten(self):
        layerShape = self.layer.get_shape()
        self.preFlatShapes.append(layerShape)
        flatSize = int(layerShape[1]) * int(layerShape[2]) * int(layerShape[3])
        self.layer = G.reshape(self.layer, [-1, flatSize])
        return flatSize

    # tools_wscale/GAN.py:461-469
    def unflatten(self):
        unflatShape = self.preFlatShapes.pop()
        self.layer = G.reshape(self.layer, [-1, int(unflatShape[1]), int(unflatShape[2]), int(unflatShape[3])])
        return self.layer

    # tools_wscale/GAN.py:438-456
    def fully_connected_layer(self, _numHidden, _act, name="full", gain=np.sqrt(2)):
        with G.variable_scope(name):
            self.layer_num += 1
            numInput = int(self.layer.get_shape()[1])
            W = self.weight_variable([numInput, _numHidden], name=name, gain=gain)
            b = self.bias_variable([_numHidden], name=name)
            self.DOFs += numInput * _numHidden + _numHidden
            self.layer = self._g.add("fc", [self.layer], (None, _numHidden), weight=W, bias=b)
            kind = _act_kind(_act)
            if kind:
                self.layer = G.activation(self.layer, kind)
            return self.layer

    # tools_wscale/GAN.py:635-638
    def concat(self, layer):
        self.layer = G.concat([self.layer, layer], axis=-1)
        return self.layer

    # tools_wscale/GAN.py:652-657
    def y(self):
        return self.layer

    def getDOFs(self):
        return self.DOFs

    # tools_wscale/GAN.py:661-678 -- He-std runtime weight scaling ("wscale"): v * float32(gain/sqrt(fan_in))
    def weight_variable(self, shape, name="w", gain=np.sqrt(2), use_he=False, in_lay=None, use_wscale=True):
        if in_lay is None:
            in_lay = np.prod(shape[:-1])
        std = gain / np.sqrt(in_lay)
        if not use_wscale:
            raise NotImplementedError("he_normal-initialised variables (use_wscale=False) are unused by the scripts")
        var = self._g.get_variable("weight", shape, "normal")
        return {"var": var, "wscale": np.float32(std)}

    # tools_wscale/GAN.py:682-683
    def bias_variable(self, shape, name="b"):
        return self._g.get_variable("bias", shape, ("const", 0.1))

    # tools_wscale/GAN.py:686-691
    def conv2d(self, x, W, stride=[1], _bias=None, _bn=None):
        if len(stride) == 2 and stride[0] != stride[1]:
            raise NotImplementedError("anisotropic strides are unused by the scripts")
        s = int(stride[0])
        xs = x.get_shape()
        kh, kw, cin, cout = W["var"].shape
        if xs[3] != cin:
            raise ValueError("conv2d: input has %d channels, filter expects %d" % (xs[3], cin))
        oh, ow = -(-xs[1] // s), -(-xs[2] // s)
        return self._g.add("conv", [x], (None, oh, ow, cout), weight=W, bias=_bias, bn=_bn, stride=s,
                           ksize=int(kh))
