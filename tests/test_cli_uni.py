"""The .uni codec against files written by the reference's own tools_wscale/uniio.writeUni (tests/golden/ref_*.uni),
the flag parser against the paramhelpers semantics, and (GPU) the drop-in command line end to end."""
import gzip
import os
import sys

import numpy as np
import pytest

import mpgan_b200  # noqa: F401
from mpgan_b200 import cli, uni

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_read_reference_written_uni_and_write_identical_bytes(tmp_path):
    g = np.load(os.path.join(GOLD, "uni.npz"))
    for name, key in (("ref_density.uni", "dens"), ("ref_velocity.uni", "vel")):
        head, data = uni.read_uni(os.path.join(GOLD, name))
        assert list(head) == ["dimX", "dimY", "dimZ", "gridType", "elementType", "bytesPerElement", "info", "dimT", "timestamp"]
        assert (head["dimX"], head["dimY"], head["dimZ"]) == (5, 4, 3) and head["timestamp"] == 1234567890123
        np.testing.assert_array_equal(data, g[key])
        out = str(tmp_path / name)
        uni.write_uni(out, head, data)
        with gzip.open(out, "rb") as a, gzip.open(os.path.join(GOLD, name), "rb") as b:
            assert a.read() == b.read()  # same container bytes as the reference writer


def test_uni_errors(tmp_path):
    p = str(tmp_path / "bad.uni")
    with gzip.open(p, "wb") as fh:
        fh.write(b"M4T3" + b"\0" * 288)
    with pytest.raises(uni.UniError):
        uni.read_uni(p)
    with pytest.raises(uni.UniError):
        uni.write_uni(p, uni.make_header((2, 2, 2)), np.zeros(7, np.float32))


def test_flag_grammar_matches_paramhelpers():
    ph = cli.Params(["prog", "simSize", "64", "UPRES", "8", "bogus", "1"])
    assert ph.get("simsize", 32) == "64" and ph.get("upRes", 4) == "8" and ph.get("tileSize", 16) == 16
    with pytest.raises(SystemExit):
        ph.check_unused()  # unknown flag aborts (paramhelpers.py:29-37)
    ph.get("bogus", 0)
    ph.check_unused()


def test_cli_requires_weights_source(tmp_path):
    with pytest.raises(SystemExit):
        cli.main(["prog", "load_model_test_1", "0", "load_model_no_1", "0", "useVelocities", "1", "simSize", "4", "tileSize", "4"])


@pytest.mark.gpu
def test_cli_end_to_end_writes_source_uni(tmp_path):
    import torch
    from mpgan_b200 import pipeline as P, synth
    L, u = 4, 4
    sim = tmp_path / "sim_1000"
    sim.mkdir()
    frames = [synth.synthetic_volume(L, seed=s) for s in (1, 2)]
    for f, x in enumerate(frames):
        uni.write_uni(str(sim / ("density_low_%04d.uni" % f)), uni.make_header((L, L, L), 1), x[..., 0:1])
        uni.write_uni(str(sim / ("velocity_low_%04d.uni" % f)), uni.make_header((L, L, L), 2), x[..., 1:4])
    flags = dict(out=1, randomInit=3, precision="fp32", packedSimPath=str(tmp_path) + "/", fromSim=1000, frame_min=0, frame_max=2,
                 simSize=L, tileSize=L, upRes=u, useVelocities=1, genUni=1, transposeAxis=0, pixelNorm=1, batchNorm=0,
                 addBicubicUpsample=1, upsampleMode=1, firstNNArch=1, velScale=1.0,
                 load_model_test_1=0, load_model_no_1=0, use_res_net1=1, add_adj_idcs1=1, startFms1=32, maxFms1=32, filterSize1=3,
                 load_model_test_2=0, load_model_no_2=0, use_res_net2=1, add_adj_idcs2=0, startFms2=32, maxFms2=32, filterSize2=5)
    argv = ["multipassGAN-out.py"]
    for k, v in flags.items():
        argv += [k, str(v)]
    assert cli.main(argv) == 0
    specs = {1: P.NetSpec(True, True, 32, 32, 3, True), 2: P.NetSpec(True, False, 32, 32, 5)}
    w = P.make_weights_out(L, 3, upRes=u, specs=specs, nets=(1, 2))
    mp = P.MultiPassOut(L, w, upRes=u, specs=specs, precision="fp32")
    for f, x in enumerate(frames):
        head, vol = uni.read_uni(str(sim / ("source_%04d.uni" % f)))
        assert (head["dimX"], head["dimY"], head["dimZ"]) == (L * u,) * 3
        np.testing.assert_array_equal(vol[..., 0], mp(x).cpu().numpy())
