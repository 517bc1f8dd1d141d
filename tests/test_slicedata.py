"""3-D volume -> 2-D training slices (mpgan_b200.slicedata) against golden vectors produced by executing the reference's
own statements of FluidDataLoader.loadFiles and its helper methods (tests/golden/make_golden.py slices)."""
import os

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import slicedata as sd

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slicedata.npz"))


@pytest.mark.parametrize("tag", ["mode2", "mode3", "mode1_tempo", "adj"])
def test_slices_from_volumes_matches_reference(tag):
    cfg = GOLD[tag + "_cfg"]
    axis, adj = int(cfg[0]), bool(cfg[2])
    sc, scy = [float(v) for v in cfg[3:7]], [float(v) for v in cfg[7:11]]
    x, y = sd.slices_from_volumes(torch.from_numpy(GOLD[tag + "_fx"]), torch.from_numpy(GOLD[tag + "_fy"]), conv_axis=axis,
                                  axis_scaling=sc, axis_scaling_y=scy, density_threshold=0.002, select_random=0.5,
                                  add_adj_idcs=adj)
    wx, wy = GOLD[tag + "_x"], GOLD[tag + "_y"]
    assert tuple(x.shape) == wx.shape and tuple(y.shape) == wy.shape and wx.shape[0] > 0
    np.testing.assert_array_equal(x.numpy(), wx)  # transposes, swaps, slice selection: pure data movement
    if all(v == 1.0 for v in scy):
        np.testing.assert_array_equal(y.numpy(), wy)
    else:  # order-1 zoom of the high-res volume along z (scipy interpolates in double precision)
        np.testing.assert_allclose(y.numpy(), wy, rtol=0, atol=2e-6)


def test_convert_slices_axis_and_channel_bookkeeping():
    """fluiddataloader.py:414-431: axis 1 -> (y,z,x) slices with vy<->vz, axis 2 -> (x,y,z) slices with vx<->vz."""
    v = torch.arange(2 * 3 * 4 * 4, dtype=torch.float32).reshape(2, 3, 4, 4)
    a1 = sd.convert_slices(v, 1)
    assert tuple(a1.shape) == (3, 2, 4, 4)
    assert torch.equal(a1[1, 0, 2], v[0, 1, 2][[0, 1, 3, 2]])
    a2 = sd.convert_slices(v, 2)
    assert tuple(a2.shape) == (4, 3, 2, 4)
    assert torch.equal(a2[3, 1, 0], v[0, 1, 3][[0, 3, 2, 1]])
    assert sd.convert_slices(v, 0) is v
    with pytest.raises(ValueError):
        sd.convert_slices(v, 3)


def test_frame_indices_match_the_fluid_data_loader():
    """Which frames a data_fraction selects: the loader's own statements (tools_wscale/fluiddataloader.py:238-244) executed for
    several index ranges (tests/golden/tempotiles.npz frac_*), incl. the ranges of the shipped 8x training command."""
    import json
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tempotiles.npz"))
    want = json.loads(str(G["frac_picked"]))
    for (lo, hi, frac), w in zip(G["frac_cases"], want):
        assert sd.frame_indices(int(lo), int(hi), float(frac)) == w, (lo, hi, frac)
    assert sd.frame_indices(0, 120, 0.08) == [0, 13, 26, 40, 53, 66, 80, 93, 106]
