"""Symbolic graph recorded by the drop-in layer API (GAN.py) instead of a TensorFlow-1.x graph.

The reference builds its generators with TF graph-mode calls (tf.placeholder, tf.reshape, tf.concat,
tf.slice, tf.image.resize_images, tf.nn.relu, tf.add, tf.variable_scope + the `GAN` class).  This
module provides the same vocabulary on a tiny IR that `engine.py` compiles into fused sm_100a
kernel launches.  Nothing here computes: tensors carry only shapes (batch dimension symbolic).
"""
import contextlib

import numpy as np


class Tensor:
    __slots__ = ("graph", "node", "shape", "name")

    def __init__(self, graph, node, shape, name=None):
        self.graph = graph
        self.node = node  # producing Node
        self.shape = tuple(shape)  # shape[0] is None (symbolic batch)
        self.name = name

    def get_shape(self):
        return self.shape

    def __add__(self, other):
        return add(self, other)

    def __repr__(self):
        return "Tensor(%s, %s)" % (self.node.op, self.shape)


class Node:
    __slots__ = ("op", "inputs", "attrs", "out", "id")

    def __init__(self, op, inputs, attrs):
        self.op = op
        self.inputs = list(inputs)
        self.attrs = attrs
        self.out = None
        self.id = -1


class Variable:
    __slots__ = ("name", "shape", "kind")

    def __init__(self, name, shape, kind):
        self.name, self.shape, self.kind = name, tuple(int(s) for s in shape), kind


class Graph:
    def __init__(self):
        self.nodes = []
        self.variables = {}  # name -> Variable (creation order preserved)
        self.scopes = []
        self.placeholders = {}

    def add(self, op, inputs, shape, **attrs):
        n = Node(op, inputs, attrs)
        n.id = len(self.nodes)
        n.out = Tensor(self, n, shape)
        self.nodes.append(n)
        return n.out

    def get_variable(self, leaf, shape, kind):
        """tf.get_variable inside the current scope; AUTO_REUSE semantics (name -> one variable)."""
        name = "/".join(self.scopes + [leaf])
        v = self.variables.get(name)
        if v is None:
            v = self.variables[name] = Variable(name, shape, kind)
        elif v.shape != tuple(int(s) for s in shape):
            raise ValueError("variable %s re-requested with shape %s (has %s)" % (name, shape, v.shape))
        return v


_default = [Graph()]


def get_default_graph():
    return _default[-1]


def reset_default_graph():
    _default[-1] = Graph()
    return _default[-1]


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    g = get_default_graph()
    g.scopes.append(name)
    try:
        yield name
    finally:
        g.scopes.pop()


def get_variable_scope():
    return "/".join(get_default_graph().scopes)


# ------------------------------------------------------------------ graph-mode ops
def placeholder(shape, name):
    """tf.placeholder(tf.float32, shape, name) -- GAN/multipassGAN-out.py:342-343."""
    g = get_default_graph()
    t = g.add("placeholder", [], shape, name=name)
    t.name = name
    g.placeholders[name] = t
    return t


def _resolve(shape, known):
    """Resolve one -1 in `shape` given the per-sample element count of the input."""
    shape = [(-1 if s is None else int(s)) for s in shape]
    assert shape[0] == -1, "only batch-leading reshapes are used by the reference graphs"
    tail = int(np.prod(shape[1:]))
    assert tail == known, "reshape %s does not match %d elements per sample" % (shape, known)
    return (None,) + tuple(shape[1:])


def reshape(x, shape):
    per = int(np.prod(x.shape[1:]))
    return x.graph.add("reshape", [x], _resolve(shape, per))


def concat(values, axis=-1):
    assert axis in (-1, 3)
    c = sum(v.shape[3] for v in values)
    s = values[0].shape
    return values[0].graph.add("concat", list(values), (None, s[1], s[2], c))


def slice_channels(x, c0, c1):
    """tf.slice(x, [0,0,0,c0], [-1,H,W,c1-c0]) -- GAN/multipassGAN-out.py:330,332."""
    s = x.shape
    return x.graph.add("slice", [x], (None, s[1], s[2], c1 - c0), c0=int(c0), c1=int(c1))


def resize_images(x, size, method=1):
    """tf.image.resize_images(x, [H, W], method): 0 bilinear, 1 nearest, 2 TF1 legacy bicubic."""
    s = x.shape
    oh, ow = int(size[0]), int(size[1])
    return x.graph.add("resize", [x], (None, oh, ow, s[3]), method=int(method))


def add(a, b):
    assert a.shape[1:] == b.shape[1:], (a.shape, b.shape)
    return a.graph.add("add", [a, b], a.shape)


def activation(x, kind):
    return x.graph.add("act", [x], x.shape, kind=kind)


def relu(x):
    return activation(x, "relu")


def lrelu(x, leak=0.2, name="lrelu"):
    """tools_wscale/GAN.py:733-737 (leak fixed at 0.2 by every caller)."""
    assert abs(leak - 0.2) < 1e-12
    return activation(x, "lrelu")


def tanh(x):
    return activation(x, "tanh")


def pixel_norm(x, epsilon=1e-8):
    assert abs(epsilon - 1e-8) < 1e-20
    return x.graph.add("pixel_norm", [x], x.shape)
