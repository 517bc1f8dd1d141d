"""Deterministic variable initialisation (random-init weights for benchmarks and parity tests).

The reference can only restore TF1 checkpoints (GAN/multipassGAN-out.py:367-386); the benchmark
harness therefore injects weights.  Each variable is seeded by (seed, crc32(name)) so its value does
not depend on creation order; values follow the reference initialisers: N(0,1) for `weight`
(tools_wscale/GAN.py:668, the He/wscale factor is applied at run time), 0.1 for `bias` (:683), and
the tf.contrib.layers.batch_norm defaults (gamma 1, beta 0, moving_mean 0, moving_variance 1).
"""
import zlib

import numpy as np


def init_variable(seed, name, shape, kind):
    if kind == "normal":
        rng = np.random.default_rng(np.random.SeedSequence([int(seed), zlib.crc32(name.encode("utf-8"))]))
        return rng.standard_normal(size=tuple(shape), dtype=np.float32)
    if kind[0] == "const":
        return np.full(tuple(shape), kind[1], dtype=np.float32)
    raise ValueError("unknown initialiser %r" % (kind,))


def init_graph_variables(graph, seed, prefix=""):
    """name -> float32 array for every variable of `graph` (keys are the reference variable names,
    i.e. what tf.train.Saver stores after stripping the `gen_N/` prefix and `:0`)."""
    return {prefix + v.name: init_variable(seed, prefix + v.name, v.shape, v.kind) for v in graph.variables.values()}


def randomize_bn_stats(weights, seed):
    """Give BN variables non-trivial values (tests): gamma/beta/moving stats away from identity."""
    out = dict(weights)
    for name in sorted(weights):
        leaf = name.rsplit("/", 1)[-1]
        rng = np.random.default_rng(np.random.SeedSequence([int(seed), zlib.crc32(name.encode("utf-8")), 7]))
        shp = weights[name].shape
        if leaf == "gamma":
            out[name] = (0.5 + rng.random(shp)).astype(np.float32)
        elif leaf == "beta":
            out[name] = (0.2 * rng.standard_normal(shp)).astype(np.float32)
        elif leaf == "moving_mean":
            out[name] = (0.3 * rng.standard_normal(shp)).astype(np.float32)
        elif leaf == "moving_variance":
            out[name] = (0.5 + rng.random(shp)).astype(np.float32)
    return out
