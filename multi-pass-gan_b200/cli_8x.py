"""Command line of `GAN/multipassGAN-8x.py` in TRAINING mode (`out 0`) for the first network (`upsamplingMode 2`, the first
call of GAN/example_run_training.py:4) on the B200 path:

    python multipassGAN-8x.py out 0 upRes 8 tileSize 16 simSize 64 use_wgan_gp 1 firstNNArch 1 upsamplingMode 2 \\
        packedSimPath ../data3d_growing/ basePath ../2ddata/ fromSim 1000 toSim 1016 frame_min 0 frame_max 120 ...

Flags: GAN/multipassGAN-8x.py:28-166, the reference's `name value` grammar (an unused flag aborts).  What runs is
training8x.Trainer8x.train: growing_gen / growing_disc / growing_disc_tempo with WGAN-GP, the staged Adam optimizers, the
generator's moving averages, the growing / blending / learning-rate schedule of the loop (:1884-2089), three-frame tiles with
semi-Lagrangian positions for the temporal critic (getTempoinput) and `model_%04d.ckpt` + `model_ema_%04d.ckpt` in
`basePath/test_%04d/` that multipassGAN-out.py restores.

Data (:250-330, 1916-1963): every growing stage trains on 2-D slices of the 3-D simulations `packedSimPath/sim_%04d/`:
inputs `density_low_%04d.uni` + `velocity_low_%04d.uni` of three consecutive frames, targets `density_low_<2|4>_%04d.uni` for
the intermediate stages and `density_high_%04d.uni` for the last one (`conv_slices`, conv_axis 0, the input z-zoomed by the
stage's factor, empty slices removed, adjacent-slice densities appended with `add_adj_idcs 1`: slicedata.py).  The reference
re-loads a stage's files at its growing event; here every stage is loaded up front and stays resident on the device.
The frames are the ones FluidDataLoader picks (slicedata.frame_indices: `data_fraction` of the index range, evenly spread;
the first stage loads max(data_fraction * 2 / currentUpres, 0.08) of [frame_min, frame_max) (:326), a later stage
`data_fraction` of the range shifted by 3 * (log2(currentUpres) - 1) frames (:1930-1937)); each index is the FIRST frame of a
triplet (multi_file_idxOff 0, 1, 2), so the files reach two frames past the last index.
Not on this command line: output mode (`out 1`: multipassGAN-out.py), the refinement networks (`upsamplingMode 1 / 3`:
Trainer8x(upsampling_mode=1) from Python), LSGAN, batch norm, dropout, pixel shuffle, 3-D data, the test / summary / PNG
side outputs of the loop.
"""
import os
import random
import sys
import time

import numpy as np

from .cli import Params

_ACCEPTED = dict(  # flags of GAN/multipassGAN-8x.py:28-166 that are read and either unused in training or fixed on this path
    numOut=200, saveOut=0, loadOut=-1, img=1, gif=0, ref=0, genModel="gen_test", discModel="disc_test", dropout=1.0,
    dropoutOutput=1.0, lambda2_f=1.0, lambda2_l1=1.0, lambda2_l2=1.0, lambda2_l3=1.0, lambda2_l4=1.0, batchSizeGen=-1,
    trainGAN=1, trainingIterations=100000, bnDecay=0.999, useVorticities=0, useFlags=0, useK_Eps_Turb=0, premadeTiles=0,
    transposeAxis=0, pretrain=0, pretrainDisc=0, pretrainGen=0, testPathStartNo=0, testInterval=100, numTests=-1,
    keepMax=3, genTestImg=-1, note="", change_velocity=0, saveMetaData=0, velScale=1.0, genUni=0, upsampleFirst=1,
    loadEmas=0, useVelInTDisc=0, lossScaling=0, outNNTestNo=17, gDrop=0, use_mb_stddev=0)


def load_stage_slices(packedSimPath, sims, frames, cu, upRes, add_adj_idcs, density_threshold, select_random, device):
    """(x [n, L, L, C*3], y [n, L*cu, L*cu, 3]) slices of frame triplets for the growing stage with data factor cu."""
    import torch
    from . import slicedata, uni
    high_name = "density_high_%04d.uni" if cu == upRes else "density_low_%i" % cu + "_%04d.uni"      # :313-316, 1935-1938
    xs, ys = [], []
    for sim in sims:
        d = os.path.join(packedSimPath, "sim_%04d" % sim)
        for f in frames:
            lows, highs = [], []
            for k in range(3):                                                                       # multi_file_idxOff 0, 1, 2
                _, dens = uni.read_uni(os.path.join(d, "density_low_%04d.uni" % (f + k)))
                _, vel = uni.read_uni(os.path.join(d, "velocity_low_%04d.uni" % (f + k)))
                _, hi = uni.read_uni(os.path.join(d, high_name % (f + k)))
                lows += [dens.astype(np.float32), vel.astype(np.float32)]
                highs.append(hi.astype(np.float32))
            fx = torch.from_numpy(np.concatenate(lows, axis=-1)).to(device)                          # [Z,Y,X,(d,vx,vy,vz) x 3]
            fy = torch.from_numpy(np.concatenate(highs, axis=-1)).to(device)                         # [Zc,Yc,Xc,3]
            x, y = slicedata.slices_from_volumes(fx, fy, conv_axis=0, axis_scaling=(cu, 1, 1, 1), axis_scaling_y=(1, 1, 1, 1),
                                                 density_threshold=density_threshold, select_random=select_random,
                                                 add_adj_idcs=add_adj_idcs)
            xs.append(x)
            ys.append(y)
    return torch.cat(xs), torch.cat(ys)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    ph = Params(argv)
    g = ph.get
    out_flag = int(g("out", 0)) > 0
    basePath = g("basePath", "../2ddata_gan/")
    randSeed = int(g("randSeed", 1))
    load_test, load_no = int(g("load_model_test", -1)), int(g("load_model_no", -1))
    simSizeLow, tileSizeLow, upRes = int(g("simSize", 64)), int(g("tileSize", 16)), int(g("upRes", 4))
    packedSimPath = g("packedSimPath", "/data/share/GANdata/2ddata_sim/")
    fromSim, toSim = int(g("fromSim", 1000)), int(g("toSim", -1))
    dataDim = int(g("dataDim", 2))
    frame_min, frame_max = int(g("frame_min", 0)), int(g("frame_max", 200))
    learning_rate, decayLR = float(g("learningRate", 0.0002)), int(g("decayLR", 0)) > 0
    beta1, beta2 = float(g("adam_beta1", 0.5)), float(g("adam_beta2", 0.999))
    weight_dld, k, k2, k_f = float(g("weight_dld", 1.0)), float(g("lambda", 1.0)), float(g("lambda2", 0.0)), float(g("lambda_f", 1.0))
    kt, kt_l = float(g("lambda_t", 1.0)), float(g("lambda_t_l2", 0.0))
    batch_size = int(g("batchSize", 128))
    batch_size_disc = int(g("batchSizeDisc", batch_size))
    discRuns, genRuns = int(g("discRuns", 1)), int(g("genRuns", 1))
    batch_norm, pixel_norm = int(g("batchNorm", 0)) > 0, int(g("pixelNorm", 1)) > 0
    useVelocities = int(g("useVelocities", 0))
    useDataAugmentation = int(g("dataAugmentation", 0))
    minScale, maxScale, rot, flip = float(g("minScale", 0.85)), float(g("maxScale", 1.15)), int(g("rot", 2)), int(g("flip", 1))
    outputInterval, saveInterval = int(g("outputInterval", 100)), int(g("saveInterval", 200))
    alwaysSave = int(g("alwaysSave", 1)) > 0
    data_fraction = float(g("data_fraction", 0.3))
    adv_flag, adv_mode = int(g("adv_flag", 1)), int(g("adv_mode", 1))
    use_spatialdisc = int(g("use_spatialdisc", 1))
    upsampling_mode, upsampled_data = int(g("upsamplingMode", 2)), int(g("upsampledData", 0))
    usePixelShuffle, addBicubic = int(g("usePixelShuffle", 0)), int(g("addBicubicUpsample", 0))
    startingIter = int(g("startingIter", 0))
    upsampleMode = int(g("upsampleMode", 1))
    stageIter, decayIter = int(g("stageIter", 25000)), int(g("decayIter", 25000))
    max_fms, start_fms, filterSize = int(g("maxFms", 256)), int(g("startFms", 512)), int(g("filterSize", 3))
    use_wgan_gp, use_res_net, use_LSGAN = int(g("use_wgan_gp", 0)), int(g("use_res_net", 0)), int(g("use_LSGAN", 0))
    first_nn_arch, add_adj_idcs = int(g("firstNNArch", 0)), int(g("add_adj_idcs", 0))
    gpu = int(str(g("gpu", 2)).split(",")[0])
    max_iters = g("maxIters", None)       # extension: stop after this many iterations (smoke runs)
    for name, default in _ACCEPTED.items():
        g(name, default)
    ph.check_unused()

    def need(cond, msg):
        if not cond:
            raise SystemExit("multipassGAN-8x: " + msg)

    need(not out_flag, "output mode (`out 1`) is multipassGAN-out.py on this path")
    need(dataDim == 2 and upsampling_mode == 2 and not upsampled_data,
         "this command line trains the first network on 2-D slices (dataDim 2, upsamplingMode 2, upsampledData 0); the "
         "refinement networks train through mpgan_b200.training8x.Trainer8x(upsampling_mode=1)")
    need(use_wgan_gp and not use_LSGAN and use_spatialdisc, "built: use_wgan_gp 1, use_LSGAN 0, use_spatialdisc 1")
    need(first_nn_arch and use_res_net and pixel_norm and not batch_norm and not usePixelShuffle and upsampleMode == 1,
         "built: firstNNArch 1, use_res_net 1, pixelNorm 1, batchNorm 0, usePixelShuffle 0, upsampleMode 1")
    need(useVelocities == 1, "useVelocities 1 is required (inputs d, vx, vy, vz)")
    need(kt_l <= 1e-6 and k2 == 0.0, "lambda_t_l2 and lambda2 terms are not built (both 0 in the shipped commands)")
    need(kt <= 1e-6 or adv_flag == 0 or adv_mode == 0, "adv_mode 1 / 2 (in-graph advection) is not built; the shipped command uses adv_mode 0")
    need(decayLR, "decayLR 0 cannot build the reference's discriminator optimizers either (one learning rate for three, :1010-1017)")
    need(rot != 2 or not useDataAugmentation, "free-angle rotation (rot 2) is not ported; the shipped command uses rot 1")
    need(upRes == 8, "the growing schedule of the script is written for upRes 8 (three stages)")
    import torch
    from . import schedule8x, tilesampler, training8x

    if toSim == -1:
        toSim = fromSim
    torch.cuda.set_device(gpu)
    device = torch.device("cuda", gpu)
    random.seed(randSeed)
    np.random.seed(randSeed)
    C = 1 + 3 + (2 if add_adj_idcs else 0)
    sched = schedule8x.GrowthSchedule(stageIter, decayIter, upRes, upsampling_mode, startingIter, decayLR)
    from . import slicedata
    # the test_%04d directory of this run (ph.getNextTestPath, :381-390)
    no = 0
    while os.path.exists(os.path.join(basePath, "test_%04d" % no)):
        no += 1
    test_path = os.path.join(basePath, "test_%04d" % no)
    os.makedirs(test_path)
    print("Called with: " + " ".join(argv))
    print("test path: " + test_path)
    stages = [u for u in (2, 4, 8) if u >= sched.initial_upres()]
    samplers = {}
    t0 = time.time()
    for cu in stages:
        if cu == stages[0]:                                                              # :326 (min_data_fraction 0.08, :247)
            frames = slicedata.frame_indices(frame_min, frame_max, max(data_fraction * 2 / cu, 0.08))
        else:                                                                            # :1930-1937 (stride = 3, :1897)
            shift = 3 * (int(round(np.log2(cu))) - 1)
            frames = slicedata.frame_indices(frame_min + shift, frame_max + shift, data_fraction)
        x, y = load_stage_slices(packedSimPath, range(fromSim, toSim + 1), frames, cu, upRes, bool(add_adj_idcs),
                                 0.005 if cu == stages[0] else 0.002, 0.4, device)       # :326 / :1937
        s = tilesampler.TileSampler(tileSizeLow, cu, densityMinimum=0.002 if cu == stages[0] else 0.01, device=device,
                                    rng=random.Random(randSeed + cu), dim_t=3)              # :301 / :1926
        s.add_data(x[:, None].cpu().numpy(), y[:, None].cpu().numpy())
        if useDataAugmentation:
            s.init_data_augmentation(rot=rot, minScale=minScale, maxScale=maxScale, flip=bool(flip),
                                     np_rng=np.random.RandomState(randSeed + cu))
        samplers[cu] = s
        print("stage %dx: %d slices of %d frame triplets (%.1f s)" % (cu, x.shape[0], len(frames) * (toSim - fromSim + 1), time.time() - t0))
    B = batch_size_disc
    tr = training8x.Trainer8x(tileSizeLow, upRes, C, start_fms, max_fms, filterSize, batch=B, learning_rate=learning_rate,
                              adam_beta1=beta1, adam_beta2=beta2, lambda_l1=k, seed=randSeed, device=gpu, upsampling_mode=2,
                              lambda_t=kt)
    if load_test >= 0:
        tr.load(os.path.join(basePath, "test_%04d" % load_test), load_no)                   # :1373-1377
        print("Model restored from test_%04d, %04d." % (load_test, load_no))
    batches = training8x.StageBatches(samplers, B, augment=bool(useDataAugmentation), tile_t=1)
    tempo = training8x.TempoBatches(samplers, B, n_t=3, dt=0.5, device=gpu, augment=bool(useDataAugmentation)) if kt > 1e-6 else None
    if tempo is not None and not adv_flag:
        inner = tempo
        tempo = lambda upres: inner(upres)[:2]                                              # adv_flag 0: frames taken as aligned
    print("\n*****TRAINING STARTED*****\n")
    hist = tr.train(batches, sched, discRuns=discRuns, genRuns=genRuns, lambda_f=k_f, add_adj_idcs=bool(add_adj_idcs),
                    save_dir=test_path, saveInterval=saveInterval, alwaysSave=alwaysSave, log=print, log_interval=outputInterval,
                    lerp_seed=randSeed, max_iters=int(max_iters) if max_iters is not None else None, tempo_batches=tempo)
    last = tr.save(test_path)
    print("Training finished after %d logged iterations; last model %04d in %s." % (len(hist), last, test_path))
    return 0
