"""Training batch source against golden vectors produced by the reference's own
TileCreator.selectRandomTiles / getRandomDatum / getRandomTile / hasMinDensity (tests/golden/make_golden.py sampler)."""
import os
import random

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import tilesampler as ts

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tilesampler.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_select_random_tiles_matches_reference_bit_for_bit(tag):
    T, L, u, nframes, dmin, part_train, seed = GOLD[tag + "_cfg"]
    part_test = {0.9: 0.1, 0.6: 0.4}[float(part_train)]
    s = ts.TileSampler(int(T), int(u), densityMinimum=float(dmin), partTrain=float(part_train), partTest=part_test,
                       rng=random.Random(int(seed)))
    s.add_data(GOLD[tag + "_low"], GOLD[tag + "_high"])
    assert s.set_borders == [int(b) for b in GOLD[tag + "_borders"]]
    for call in range(3):  # same generator state carried across calls, like the training loop
        low, high = s.select_random_tiles(5, is_training=True)
        np.testing.assert_array_equal(low.numpy(), GOLD[tag + "_train_low"][call])
        np.testing.assert_array_equal(high.numpy(), GOLD[tag + "_train_high"][call])
    low, high = s.select_random_tiles(4, is_training=False)
    np.testing.assert_array_equal(low.numpy(), GOLD[tag + "_test_low"])
    np.testing.assert_array_equal(high.numpy(), GOLD[tag + "_test_high"])


def test_density_rejection_retries_and_rows_layout():
    """The fixture is sparse with an empty half: nearly all picks must pass the density test, the rows are the C-order
    flattening of the tiles, and a frame without density stops after 19 tries per tile."""
    T, L, u = 8, 24, 4
    low, high = GOLD["a_low"], GOLD["a_high"]
    s = ts.TileSampler(T, u, densityMinimum=0.02, rng=random.Random(3))
    s.add_data(low, high)
    picks = s.select_offsets(64)
    dens = [float(low[f, 0, oy:oy + T, ox:ox + T, 0].sum(dtype=np.float64)) for f, oy, ox, _ in picks]
    assert sum(d >= 0.02 * T * T for d in dens) >= 60
    x, y = s.batch_rows(6)
    assert tuple(x.shape) == (6, T * T * 4) and tuple(y.shape) == (6, (T * u) ** 2)
    s2 = ts.TileSampler(T, u, densityMinimum=0.02, rng=random.Random(9))
    s2.add_data(low, high)
    p = s2.select_offsets(2)
    lo, hi = s2.gather(p)
    f, oy, ox, _ = p[1]
    np.testing.assert_array_equal(lo[1, 0].numpy(), low[f, 0, oy:oy + T, ox:ox + T])
    np.testing.assert_array_equal(hi[1, 0].numpy(), high[f, 0, oy * u:(oy + T) * u, ox * u:(ox + T) * u])
    calls = []

    class Counting(random.Random):
        def randrange(self, a, b=None):
            calls.append((a, b))
            return super().randrange(a, b)

    s3 = ts.TileSampler(T, u, densityMinimum=0.02, rng=Counting(1))
    s3.add_data(np.zeros_like(low), high)
    s3.select_offsets(1)
    assert len(calls) == 1 + 19 * 3  # one frame pick + 19 tries of (z, y, x) offsets (i < 20, tilecreator_t.py:624)


def test_errors_mirror_the_reference():
    s = ts.TileSampler(8, 4)
    with pytest.raises(ts.TileSamplerError):
        s.add_data(np.zeros((2, 1, 24, 24, 4), np.float32), np.zeros((3, 1, 96, 96, 1), np.float32))
    with pytest.raises(ts.TileSamplerError):
        s.add_data(np.zeros((2, 1, 24, 24, 4), np.float32), np.zeros((2, 1, 48, 48, 1), np.float32))
    s.add_data(np.ones((1, 1, 24, 24, 4), np.float32), np.ones((1, 1, 96, 96, 1), np.float32))
    assert s.set_borders == [0, 0, 1]  # int(1 * 0.9) = 0 training frames
    with pytest.raises(ts.TileSamplerError):
        s.select_offsets(1)


@pytest.mark.gpu
def test_device_resident_sampler_feeds_the_trainer():
    """Same decisions on the GPU: frames resident in HBM, batches gathered there and consumed by Trainer4x without a
    host round trip."""
    from mpgan_b200 import training as T_
    low, high = GOLD["a_low"], GOLD["a_high"]
    cpu = ts.TileSampler(8, 4, densityMinimum=0.02, rng=random.Random(5))
    gpu = ts.TileSampler(8, 4, densityMinimum=0.02, rng=random.Random(5), device="cuda")
    cpu.add_data(low, high)
    gpu.add_data(low, high)
    xc, yc = cpu.batch_rows(4)
    xg, yg = gpu.batch_rows(4)
    assert xg.is_cuda and torch.equal(xg.cpu(), xc) and torch.equal(yg.cpu(), yc)
    tr = T_.Trainer4x(8, 4, 4, seed=2)
    out = tr.iteration([(xg, yg)], [(xg, yg)])
    assert np.isfinite(out["gen_loss_complete"]) and np.isfinite(out["disc_loss"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "t3"])
def test_augmented_tiles_match_reference_generate_tile(tag):
    """selectRandomTiles(augment=True) -> generateTile (scaling, second cut, rot90, flip, velocity fix-ups) against the
    reference's own methods (tests/golden/tileaugment.npz): identical decisions (Python random + numpy RandomState call
    sequences), data movement exact, the order-1 zoom within fp32 rounding of scipy's double-precision interpolation."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tileaugment.npz"))
    T, L, u, nframes, dmin, smin, smax, rot, flip, seed_py, seed_np = gold[tag + "_cfg"]
    # "t3": three-frame tiles (TileCreator(dim_t=3), tile_t=3): per-frame velocity fix-ups, rot90 DOES rotate the vectors there
    s = ts.TileSampler(int(T), int(u), densityMinimum=float(dmin), rng=random.Random(int(seed_py)), dim_t=3 if tag == "t3" else 1)
    s.add_data(gold[tag + "_low"], gold[tag + "_high"])
    s.init_data_augmentation(rot=int(rot), minScale=float(smin), maxScale=float(smax), flip=bool(flip),
                             np_rng=np.random.RandomState(int(seed_np)))
    exact = float(smin) == 1.0 and float(smax) == 1.0
    for call in range(3):
        low, high = s.select_random_tiles(6, is_training=True, augment=True)
        wl, wh = gold[tag + "_aug_low"][call], gold[tag + "_aug_high"][call]
        assert tuple(low.shape) == wl.shape and tuple(high.shape) == wh.shape
        if exact:  # no interpolation: cut + rot90 + flip + sign changes are pure data movement
            np.testing.assert_array_equal(low.numpy(), wl)
            np.testing.assert_array_equal(high.numpy(), wh)
        else:
            np.testing.assert_allclose(low.numpy(), wl, rtol=0, atol=3e-6)
            np.testing.assert_allclose(high.numpy(), wh, rtol=0, atol=3e-6)


@pytest.mark.parametrize("key,aug", [("single", False), ("single_aug", True)])
def test_single_frame_tiles_from_three_frame_data_match_the_reference(key, aug):
    """getinput of the 8x trainer on TileCreator(dim_t=3) data: selectRandomTiles with the reference's default tile_t = 1 first
    draws which frame of the sequence the tile comes from (never the last one, getRandomDatum :548-560), the density test and
    the augmentation then see that frame only (single-frame rule: rot90 leaves the velocity vectors alone)."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tileaugment.npz"))
    T, L, u, nframes, dmin, smin, smax, rot, flip, _, _ = gold["t3_cfg"]
    seed_py, seed_np = gold["t3_single_seeds"]
    s = ts.TileSampler(int(T), int(u), densityMinimum=float(dmin), rng=random.Random(int(seed_py)), dim_t=3)
    s.add_data(gold["t3_low"], gold["t3_high"])
    s.init_data_augmentation(rot=int(rot), minScale=float(smin), maxScale=float(smax), flip=bool(flip),
                             np_rng=np.random.RandomState(int(seed_np)))
    for call in range(3):
        low, high = s.select_random_tiles(6, is_training=True, augment=aug, tile_t=1)
        wl, wh = gold["t3_%s_low" % key][call], gold["t3_%s_high" % key][call]
        assert tuple(low.shape) == wl.shape and tuple(high.shape) == wh.shape
        np.testing.assert_allclose(low.numpy(), wl, rtol=0, atol=0 if not aug else 3e-6)
        np.testing.assert_allclose(high.numpy(), wh, rtol=0, atol=0 if not aug else 3e-6)
    with pytest.raises(ts.TileSamplerError):
        s.select_random_tiles(2, tile_t=4)


def test_augmentation_free_rotation_is_refused():
    s = ts.TileSampler(8, 4)
    with pytest.raises(NotImplementedError):
        s.init_data_augmentation(rot=2)
