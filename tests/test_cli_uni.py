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


def test_cli_4x_flag_handling():
    """GAN/multipassGAN-4x.py command line: every reference flag is accepted, unknown flags abort, and runs outside the
    two shipped output-mode calls (GAN/example_run_output.py:6,8) are refused with a message instead of guessed at."""
    from mpgan_b200 import cli_4x
    base = ["multipassGAN-4x.py", "useVelocities", "1", "simSize", "8", "tileSize", "8"]
    with pytest.raises(SystemExit):  # unknown flag
        cli_4x.main(base + ["out", "1", "noSuchFlag", "3"])
    with pytest.raises(SystemExit) as e:  # training mode
        cli_4x.main(base + ["out", "0"])
    assert "Trainer4x" in str(e.value)
    with pytest.raises(SystemExit) as e:  # third network
        cli_4x.main(base + ["out", "1", "upsamplingMode", "3", "upsampledData", "1", "randomInit", "1"])
    assert "upsamplingMode" in str(e.value)
    with pytest.raises(SystemExit):  # tiles smaller than the frame
        cli_4x.main(["multipassGAN-4x.py", "useVelocities", "1", "simSize", "8", "tileSize", "4", "out", "1", "randomInit", "1"])


@pytest.mark.gpu
def test_cli_4x_two_runs_hand_over_uni_like_the_reference_recipe(tmp_path):
    """The benchmarked recipe through its own entry point: two `multipassGAN-4x.py out 1` runs (upsamplingMode 2, then
    upsamplingMode 1 upsampledData 1) with the .uni hand-over in between (GAN/example_run_output.py:6,8) must give exactly
    the volume of the device-resident two-pass pipeline with the same weights."""
    from mpgan_b200 import cli_4x, pipeline as P, synth, graph as G, networks as N, weights as W
    L, u = 8, 4
    S = L * u
    sim = tmp_path / "sim_1005"
    sim.mkdir()
    x = synth.synthetic_volume(L, seed=5)
    uni.write_uni(str(sim / "density_low_0110.uni"), uni.make_header((L, L, L), 1), x[..., 0:1])
    uni.write_uni(str(sim / "velocity_low_0110.uni"), uni.make_header((L, L, L), 2), x[..., 1:4])
    # the shipped command lines (training flags included: they are accepted and ignored in output mode)
    common = ("randSeed 174213111 upRes 4 startIndex 0 out 1 pretrain 0 pretrainDisc 0 tileSize %d simSize %d lambda 5.0 lambda2 0.00001 "
              "discRuns 2 genRuns 2 alwaysSave 1 fromSim 1005 toSim 1005 outputInterval 200 genTestImg 1 dropout 0.5 dataDim 2 "
              "batchSize 16 useVelocities 1 useVorticities 0 useK_Eps_Turb 0 useFlags 0 gif 0 genModel gen_resnet discModel disc_binclass "
              "packedSimPath %s/ lambda_t 1.0 lambda_t_l2 0.0 frame_min 110 frame_max 111 data_fraction 0.01 adv_flag 1 "
              "dataAugmentation 1 premadeTiles 0 rot 1 sliceMode 1 genUni 1 interpMode 1 velScale 1.0 precision fp32" % (L, L, tmp_path)).split()
    assert cli_4x.main(["multipassGAN-4x.py"] + common + "upsamplingMode 2 upsampledData 0 randomInit 31".split()) == 0
    head, first = uni.read_uni(str(sim / "density_low_2x2_0110.uni"))
    assert (head["dimX"], head["dimY"], head["dimZ"]) == (S, S, S)
    assert cli_4x.main(["multipassGAN-4x.py"] + common + "upsamplingMode 1 upsampledData 1 randomInit 32".split()) == 0
    _, second = uni.read_uni(str(sim / "density_low_1x1_0110.uni"))

    def weights_for(mode, seed):
        G.reset_default_graph()
        cfg = N.config_4x(L, upRes=u, upsampling_mode=mode)
        N.gen_resnet(G.placeholder([None, (L * L if mode == 2 else S * S) * 4], "x"), cfg)
        return W.init_graph_variables(G.get_default_graph(), seed)

    mp = P.MultiPass4x(L, weights_for(2, 31), weights_for(1, 32), upRes=u, precision="fp32")
    np.testing.assert_array_equal(first[..., 0], mp.pass1_only(x).cpu().numpy())
    np.testing.assert_array_equal(second[..., 0], mp(x).cpu().numpy())


# the flags of the reference's first-network training command (GAN/example_run_training.py:4), sizes shrunk for the test
_TRAIN_8X = ("randSeed 16131119 upRes 8 use_res_net 1 batchNorm 0 pixelNorm 1 out 0 pretrain 0 pretrainDisc 0 tileSize 4 simSize 8 "
             "use_LSGAN 0 use_wgan_gp 1 lambda 1.0 lambda2 0.0 discRuns 1 genRuns 1 alwaysSave 1 fromSim 1000 toSim 1000 outputInterval 2 "
             "genTestImg 1 dropout 0.5 dataDim 2 batchSize 6 useVelocities 1 useVorticities 0 useK_Eps_Turb 0 useFlags 0 gif 0 "
             "genModel gen_resnet discModel disc_binclass lambda_t 1.0 lambda_t_l2 0.0 frame_max 3 frame_min 0 data_fraction 1.0 "
             "adv_flag 1 adv_mode 0 dataAugmentation 1 premadeTiles 0 rot 1 minScale 0.85 maxScale 1.15 flip 1 decayLR 1 adam_beta1 0.0 "
             "adam_beta2 0.99 learningRate 0.0001 lossScaling 1 stageIter 1 decayIter 1 maxFms 32 startFms 32 filterSize 3 upsamplingMode 2 "
             "upsampledData 0 discRuns 1 load_model_test -1 load_model_no -1 firstNNArch 1 add_adj_idcs 1 usePixelShuffle 0 "
             "addBicubicUpsample 1 startingIter 0 useVelInTDisc 0 upsampleMode 1 gpu 0")


def test_cli_8x_flag_checks(tmp_path):
    from mpgan_b200 import cli_8x
    base = ["multipassGAN-8x.py"] + _TRAIN_8X.split() + ["packedSimPath", str(tmp_path) + "/", "basePath", str(tmp_path) + "/"]
    with pytest.raises(SystemExit):
        cli_8x.main(base + ["bogusFlag", "1"])                       # unused flag aborts (paramhelpers)
    for k, v in (("out", "1"), ("upsamplingMode", "1"), ("use_wgan_gp", "0"), ("adv_mode", "2"), ("decayLR", "0")):
        argv = list(base)
        for i in [j for j in range(1, len(argv), 2) if argv[j] == k]:
            argv[i + 1] = v                                          # (some flags appear twice in the shipped command)
        with pytest.raises(SystemExit) as e:
            cli_8x.main(argv)
        assert "multipassGAN-8x" in str(e.value), (k, e.value)


@pytest.mark.gpu
def test_cli_8x_trains_from_uni_simulations_and_the_result_applies(tmp_path):
    """`python multipassGAN-8x.py out 0 ...` with the flags of the reference's first-network training command on a tiny
    synthetic simulation: three-frame slices of every growing stage are loaded from .uni files, the whole (tiny) schedule
    runs with the spatial and the temporal critic, and the saved moving-average checkpoint restores into multipassGAN-out.py."""
    import torch
    from mpgan_b200 import cli_8x, synth, tfckpt
    L = 8
    sim = tmp_path / "data" / "sim_1000"
    sim.mkdir(parents=True)
    rng = np.random.default_rng(3)
    # frames: stage 2x loads the triplets starting at 0, 1, 2; the 4x / 8x stages the index range shifted by 3 / 6 (:1930-1937),
    # every triplet reaches two frames further -> files 0 .. 10
    for f in range(11):
        x = synth.synthetic_volume(L, seed=10 + f)
        x[..., 0] = np.maximum(x[..., 0], 0.05)                       # keep every slice above the density threshold
        uni.write_uni(str(sim / ("density_low_%04d.uni" % f)), uni.make_header((L, L, L), 1), x[..., 0:1])
        uni.write_uni(str(sim / ("velocity_low_%04d.uni" % f)), uni.make_header((L, L, L), 2), x[..., 1:4])
        for cu, name in ((2, "density_low_2_%04d.uni"), (4, "density_low_4_%04d.uni"), (8, "density_high_%04d.uni")):
            hi = rng.random((L * cu, L * cu, L * cu, 1), dtype=np.float32)
            uni.write_uni(str(sim / (name % f)), uni.make_header((L * cu,) * 3, 1), hi)
    base = tmp_path / "runs"
    base.mkdir()
    argv = ["multipassGAN-8x.py"] + _TRAIN_8X.split() + ["packedSimPath", str(tmp_path / "data") + "/", "basePath", str(base) + "/"]
    assert cli_8x.main(argv) == 0
    test_dir = base / "test_0000"
    saved = sorted(p.name for p in test_dir.iterdir() if p.name.endswith(".index"))
    assert "model_0000.ckpt.index" in saved and "model_ema_0000.ckpt.index" in saved and len(saved) >= 6   # 2 growing events + final
    last = max(int(n[len("model_ema_"):len("model_ema_") + 4]) for n in saved if n.startswith("model_ema_"))
    got = tfckpt.read_checkpoint(str(test_dir / ("model_ema_%04d.ckpt" % last)), verify_data=True)
    assert any(k.startswith("generator/genBlock8/") for k in got) and any(k.startswith("tempo-disc/") for k in got)
    assert all(np.isfinite(v).all() for v in got.values())
    # apply the trained first network through the out.py command line (weights restored from the run's directory)
    flags = dict(out=1, packedSimPath=str(tmp_path / "data") + "/", basePath=str(base) + "/", fromSim=1000, frame_min=0, frame_max=1,
                 simSize=L, tileSize=L, upRes=8, useVelocities=1, genUni=1, transposeAxis=0, pixelNorm=1, batchNorm=0,
                 addBicubicUpsample=1, upsampleMode=1, firstNNArch=1, velScale=1.0, loadEmas=1, precision="fp32",
                 load_model_test_1=0, load_model_no_1=last, use_res_net1=1, add_adj_idcs1=1, startFms1=32, maxFms1=32, filterSize1=3,
                 load_model_test_2=-1, load_model_no_2=-1, load_model_test_3=-1, load_model_no_3=-1)
    argv = ["multipassGAN-out.py"]
    for k, v in flags.items():
        argv += [k, str(v)]
    assert cli.main(argv) == 0
    head, vol = uni.read_uni(str(sim / "source_0000.uni"))
    assert (head["dimX"], head["dimY"], head["dimZ"]) == (64, 64, 64) and np.isfinite(vol).all()


def test_cli_8x_stage_loading_builds_three_frame_slices(tmp_path):
    """load_stage_slices on the CPU: frame triplets (multi_file_idxOff 0, 1, 2) of one simulation -> slices with the channel
    groups (d, vx, vy, vz, d(z-1), d(z+1)) x 3 frames and the stage's targets x 3 frames, the input z-zoomed by the stage factor."""
    import torch
    from mpgan_b200 import cli_8x, slicedata, synth
    L, cu = 4, 2
    sim = tmp_path / "sim_1000"
    sim.mkdir()
    rng = np.random.default_rng(8)
    vols, highs = [], []
    for f in range(4):
        x = synth.synthetic_volume(L, seed=30 + f)
        x[..., 0] = np.maximum(x[..., 0], 0.05)
        hi = rng.random((L * cu,) * 3 + (1,), dtype=np.float32)
        uni.write_uni(str(sim / ("density_low_%04d.uni" % f)), uni.make_header((L, L, L), 1), x[..., 0:1])
        uni.write_uni(str(sim / ("velocity_low_%04d.uni" % f)), uni.make_header((L, L, L), 2), x[..., 1:4])
        uni.write_uni(str(sim / ("density_low_2_%04d.uni" % f)), uni.make_header((L * cu,) * 3, 1), hi)
        vols.append(x)
        highs.append(hi)
    x, y = cli_8x.load_stage_slices(str(tmp_path) + "/", [1000], [0, 1], cu, 8, True, 0.005, 1.0, "cpu")
    assert tuple(x.shape) == (2 * L * cu, L, L, 18) and tuple(y.shape) == (2 * L * cu, L * cu, L * cu, 3)
    for t0 in (0, 1):                                   # the two triplets, L * cu slices each
        xs, ys = x[t0 * L * cu:(t0 + 1) * L * cu], y[t0 * L * cu:(t0 + 1) * L * cu]
        for k in range(3):
            want = slicedata.zoom_linear(torch.from_numpy(vols[t0 + k]), (cu, 1, 1, 1))
            assert torch.allclose(xs[..., 6 * k:6 * k + 4], want, atol=1e-6)
            assert torch.equal(xs[1:, :, :, 6 * k + 4], xs[:-1, :, :, 6 * k]) and float(xs[0, :, :, 6 * k + 4].abs().max()) == 0.0
            assert torch.equal(ys[..., k], torch.from_numpy(highs[t0 + k][..., 0]))
