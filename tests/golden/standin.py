"""Deterministic stand-in "networks" for the pipeline golden vectors (test infrastructure).

The reference's volume pipeline (generate3DUniForNewNetwork) only touches its generators through
`sess.run(sampler, feed_dict)` on flat rows, so the axis / channel / batching bookkeeping can be pinned
with any deterministic row function.  These are asymmetric in (row, col) and weight every input
channel differently, so a transposed slice, a swapped velocity channel or a shifted batch changes the
result.  Pure numpy, float32 like the reference feeds.
"""
import numpy as np

COEF_IN = np.array([1.0, 0.3, -0.2, 0.5, 0.7, -0.4], dtype=np.float32)
COEF_REF = np.array([0.6, -0.35, 0.25, 0.45], dtype=np.float32)


def ramp(S, a, b):
    yy, xx = np.mgrid[0:S, 0:S]
    return (1.0 + a * yy + b * xx).astype(np.float32)


def net_first(rows, L, u, cin):
    """rows: flat floats holding [B, L, L, cin] -> [B, (L*u)^2]."""
    a = np.asarray(rows, dtype=np.float32).reshape(-1, L, L, cin)
    s = (a * COEF_IN[:cin]).sum(axis=-1, dtype=np.float32)
    up = np.repeat(np.repeat(s, u, axis=1), u, axis=2) * ramp(L * u, 0.03, 0.001)
    return np.abs(up).reshape(a.shape[0], -1).astype(np.float32)


def net_refine(xrows, yrows, L, u, tag):
    """xrows: [B, L*L*4] low-res fields, yrows: [B, S*S] previous-pass density -> [B, S*S]."""
    S = L * u
    a = np.asarray(xrows, dtype=np.float32).reshape(-1, L, L, 4)
    s = (a * COEF_REF).sum(axis=-1, dtype=np.float32)
    up = np.repeat(np.repeat(s, u, axis=1), u, axis=2)
    y = np.asarray(yrows, dtype=np.float32).reshape(-1, S, S)
    r = ramp(S, 0.002 * tag, 0.017 * tag)
    return (0.5 * y + 0.25 * np.abs(up) * r).reshape(a.shape[0], -1).astype(np.float32)


def net_fullres(rows, S, cin=4):
    """4x passes 2/3 (upsamplingMode 1/3): rows hold [B, S, S, cin] full-res inputs -> [B, S*S]."""
    a = np.asarray(rows, dtype=np.float32).reshape(-1, S, S, cin)
    s = (a * COEF_IN[:cin]).sum(axis=-1, dtype=np.float32) * ramp(S, 0.011, 0.0007)
    return np.abs(s).reshape(a.shape[0], -1).astype(np.float32)
