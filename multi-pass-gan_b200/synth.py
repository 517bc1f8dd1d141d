"""Synthetic low-res simulation frames (there is no network access for the mantaflow data sets of
datagen/gen_sim_grow_slices_data.py).  Shape and statistics follow SURVEY §8d: `[Z, Y, X, 4]` float32 =
(density, vx, vy, vz) like `FluidDataLoader` returns for `density_low` + `velocity_low`
(tools_wscale/fluiddataloader.py:400-409, tools_wscale/uniio.py:40-44): density in [0,1] as a clipped sum
of ~18 Gaussian blobs (the 18 noise sources of datagen/gen_sim_grow_slices_data.py:232-300) with a large
fraction of exact zeros, velocities N(0, 0.5^2) smoothed with a 3-voxel box filter.
"""
import numpy as np


def _box3(a):
    out = a.copy()
    for ax in range(3):
        out = (np.roll(out, 1, ax) + out + np.roll(out, -1, ax)) / 3.0
    return out


def synthetic_volume(L, seed=1, vel_scale=1.0):
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), 0x5EED]))
    z, y, x = np.meshgrid(np.arange(L), np.arange(L), np.arange(L), indexing="ij")
    dens = np.zeros((L, L, L), dtype=np.float64)
    for _ in range(18):
        c = rng.uniform(0.15, 0.85, size=3) * L
        s = rng.uniform(0.04, 0.12) * L
        a = rng.uniform(0.3, 1.0)
        dens += a * np.exp(-((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2) / (2 * s * s))
    dens = np.clip(dens - 0.35, 0.0, 1.0)  # carve out exact zeros like empty air
    vol = np.empty((L, L, L, 4), dtype=np.float32)
    vol[..., 0] = dens
    for c in range(3):
        vol[..., 1 + c] = _box3(rng.normal(0.0, 0.5, size=(L, L, L))) * vel_scale
    return vol
