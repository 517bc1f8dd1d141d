"""Drop-in command line of `GAN/multipassGAN-4x.py` in output mode (`out 1`) on the B200 path: the entry point of the
benchmarked 4x recipe, which the reference runs as TWO processes that hand the volume over as a .uni file
(GAN/example_run_output.py:6,8; output loop GAN/multipassGAN-4x.py:1634-1646 -> generate3DUniForNewNetwork :1090-1169):

    python multipassGAN-4x.py out 1 upsamplingMode 2 upsampledData 0 genUni 1 useVelocities 1 simSize 128 tileSize 128 ...
        reads  packedSimPath/sim_%04d/density_low_%04d.uni + velocity_low_%04d.uni
        writes packedSimPath/sim_%04d/density_low_2x2_%04d.uni              (first generator: xy slices along z)
    python multipassGAN-4x.py out 1 upsamplingMode 1 upsampledData 1 genUni 1 ...
        reads  velocity_low_%04d.uni + density_low_2x2_%04d.uni
        writes packedSimPath/sim_%04d/density_low_1x1_%04d.uni              (second generator: (z,y) slices along x)

Flags: GAN/multipassGAN-4x.py:32-144, same `name value` grammar as the reference's paramhelpers (case-insensitive
names, an unused flag aborts with exit code 1). Every flag of the reference is accepted; the training flags do not
influence output mode. Weights: `load_model_test` / `load_model_no` restore basePath/test_%04d/model_%04d.ckpt
(TF checkpoint V2 read by tfckpt.py; variable names generator/g_cA0/weight ... as saved by the script's tf.train.Saver),
or the extensions `randomInit <seed>` / `weightsNpz <file>` shared with multipassGAN-out.py.
Not on this path: training (`out 0`: use training.Trainer4x), upsamplingMode 0 / 3 (a third, differently shaped
network), 3-D data (`dataDim 3`), PNG previews (scipy.misc.imsave is gone from scipy).
"""
import os
import sys
import time

import numpy as np

from .cli import Params

# every flag of GAN/multipassGAN-4x.py:32-144 with its default (accepted; most only matter for training)
_REFERENCE_FLAGS = dict(
    toSim=-1, dataDim=2, numOut=200, saveOut=0, loadOut=-1, img=1, gif=0, ref=0, genModel="gen_test", discModel="disc_test",
    learningRate=0.0002, decayLR=0, dropout=1.0, dropoutOutput=1.0, adam_beta1=0.5, weight_dld=1.0, lambda2=0.0,
    lambda_f=1.0, lambda2_f=1.0, lambda2_l1=1.0, lambda2_l2=1.0, lambda2_l3=1.0, lambda2_l4=1.0, lambda_t=1.0, lambda_t_l2=0.0,
    batchSize=128, batchSizeDisc=128, batchSizeGen=128, trainGAN=1, trainingEpochs=100000, discRuns=1, genRuns=1,
    bnDecay=0.999, useVorticities=0, useFlags=0, useK_Eps_Turb=0, premadeTiles=0, cropOverlap=0, dataAugmentation=0,
    minScale=0.85, maxScale=1.15, rot=2, flip=1, pretrain=0, pretrainDisc=0, pretrainGen=0, testPathStartNo=0, testInterval=100,
    numTests=10, outputInterval=100, saveInterval=200, alwaysSave=1, keepMax=3, genTestImg=-1, note="", data_fraction=0.3,
    adv_flag=1, change_velocity=0, saveMetaData=0, use_spatialdisc=1, clamping=1, simLowLength=64, simLowWidth=64,
    simLowHeight=64, overlappedpixel=3, startIndex=0, useAvgDepool=0, avgMode=0, sliceMode=0, interpMode=1, setVelZero=0,
    trainingIterations=0, gpu="0")
_REFERENCE_FLAGS["lambda"] = 1.0


def main(argv=None):
    argv = sys.argv if argv is None else argv
    ph = Params(argv)
    g = ph.get
    out_flag = int(g("out", 0)) > 0
    basePath = g("basePath", "../2ddata_gan/")
    randSeed = int(g("randSeed", 1))  # noqa: F841 - output mode draws no random numbers
    load_test, load_no = int(g("load_model_test", -1)), int(g("load_model_no", -1))
    simSizeLow, tileSizeLow, upRes = int(g("simSize", 64)), int(g("tileSize", 16)), int(g("upRes", 4))
    packedSimPath = g("packedSimPath", "/data/share/GANdata/2ddata_sim/")
    fromSim = int(g("fromSim", 1000))
    frame_min, frame_max = int(g("frame_min", 0)), int(g("frame_max", 200))
    batch_norm = int(g("batchNorm", 1)) > 0
    useVelocities = int(g("useVelocities", 0))
    velScale = float(g("velScale", 1.0))
    upsampling_mode = int(g("upsamplingMode", 2))
    upsampled_data = int(g("upsampledData", 0)) > 0
    genUni = int(g("genUni", 0)) > 0
    upsampleFirst = int(g("upsampleFirst", 1)) > 0
    vals = {k: g(k, v) for k, v in _REFERENCE_FLAGS.items()}
    random_init = g("randomInit", None)   # extension
    weights_npz = g("weightsNpz", None)   # extension
    precision = g("precision", "fp16")    # extension: fp16 | bf16 | fp32
    range_check = int(g("rangeCheck", 0)) != 0  # extension: validation mode (saturated 16-bit activations abort)
    ph.check_unused()
    if not out_flag:
        raise SystemExit("multipassGAN-4x: only output mode (`out 1`) runs on this command line; the training loop of "
                         "GAN/multipassGAN-4x.py:1316-1397 is mpgan_b200.training.Trainer4x (README.md)")
    if int(vals["dataDim"]) != 2:
        raise SystemExit("multipassGAN-4x: dataDim 3 is outside the accelerated path")
    if tileSizeLow != simSizeLow:
        raise SystemExit("multipassGAN-4x: output mode slices whole frames (tileSize must equal simSize, as in "
                         "GAN/example_run_output.py:6,8)")
    if not useVelocities:
        raise SystemExit("multipassGAN-4x: the shipped generators take (density, vx, vy, vz): useVelocities 1 is required")
    if not ((upsampling_mode == 2 and not upsampled_data and upsampleFirst) or (upsampling_mode == 1 and upsampled_data)):
        raise SystemExit("multipassGAN-4x: supported runs are `upsamplingMode 2 upsampledData 0` (first generator) and "
                         "`upsamplingMode 1 upsampledData 1` (second generator), the two calls of GAN/example_run_output.py:6,8")
    first = upsampling_mode == 2
    import torch
    from . import graph as G, networks as N, pipeline as P, uni, weights as W

    gpu = int(str(vals["gpu"]).split(",")[0])
    if torch.cuda.is_available():
        torch.cuda.set_device(gpu)
    L, S = simSizeLow, simSizeLow * upRes
    # the variables this run's graph asks for (names / shapes of gen_resnet, GAN/multipassGAN-4x.py:528-569)
    G.reset_default_graph()
    cfg = N.config_4x(L, upRes=upRes, upsampling_mode=upsampling_mode, batch_norm=batch_norm)
    N.gen_resnet(G.placeholder([None, (L * L if first else S * S) * 4], "x"), cfg)
    want = {v.name: tuple(v.shape) for v in G.get_default_graph().variables.values()}
    if weights_npz:
        arch = np.load(weights_npz)
        weights = {k: arch[k] for k in want}
    elif random_init is not None:
        weights = W.init_graph_variables(G.get_default_graph(), int(random_init))
    else:
        from . import tfckpt
        prefix = os.path.join(basePath, "test_%04d" % load_test, "model_%04d.ckpt" % load_no)
        if load_test < 0 or not os.path.exists(prefix + ".index"):
            raise SystemExit("multipassGAN-4x: checkpoint %s.index not found; pass `randomInit <seed>` or `weightsNpz <file>` "
                             "to run without a trained model" % prefix)
        got = tfckpt.read_checkpoint(prefix, names=sorted(want), verify_data=True)
        weights = {k: np.asarray(got[k], dtype=np.float32) for k in want}
        print("Model restored from %s." % prefix)
    for k, shp in want.items():
        if tuple(weights[k].shape) != shp:
            raise SystemExit("multipassGAN-4x: '%s' has shape %s, the graph built from the flags needs %s" % (k, weights[k].shape, shp))
    mp = P.MultiPass4x(L, weights if first else None, None if first else weights, upRes=upRes, precision=precision,
                       velScale=velScale, batch_norm=batch_norm, device=gpu, threshold=P.THRESHOLD if genUni else 0.0,
                       range_check=range_check)
    sim_path = os.path.join(packedSimPath, "sim_%04d" % fromSim)
    out_name = "density_low_2x2_%04d.uni" if first else "density_low_1x1_%04d.uni"
    print("*****OUTPUT ONLY*****")
    for f in range(frame_min, frame_max):
        head, dens = uni.read_uni(os.path.join(sim_path, "density_low_%04d.uni" % f))
        _, vel = uni.read_uni(os.path.join(sim_path, "velocity_low_%04d.uni" % f))
        x = np.concatenate([dens.astype(np.float32), vel.astype(np.float32)], axis=-1)  # [Z,Y,X,(d,vx,vy,vz)]
        t0 = time.time()
        if first:
            vol = mp.pass1_only(x)
        else:
            _, d2 = uni.read_uni(os.path.join(sim_path, "density_low_2x2_%04d.uni" % f))
            vol = mp.pass2_only(x, d2.reshape(S, S, S))
        host = vol.cpu().numpy() if genUni else None
        torch.cuda.synchronize()
        print(time.time() - t0)
        if genUni:
            head = dict(head)
            head["dimX"] = head["dimY"] = head["dimZ"] = S
            uni.write_uni(os.path.join(sim_path, out_name % f), head, host)  # :1158-1168
        print("")
    print("Test finished, %d frames written to %s." % (frame_max - frame_min, sim_path))
    return 0
