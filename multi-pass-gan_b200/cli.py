"""Drop-in command line of `GAN/multipassGAN-out.py` (the apply-model entry point) on the B200 path.

    python multipassGAN-out.py  key value  key value ...        (flags: GAN/multipassGAN-out.py:28-99, App. E)

Same flag grammar as the reference's paramhelpers (`name value` pairs, names case-insensitive, values strings,
an unknown / unused flag aborts with exit code 1: tools_wscale/paramhelpers.py:16-37), same input files
(`packedSimPath/sim_%04d/density_low_%04d.uni` + `velocity_low_%04d.uni`, frames [frame_min, frame_max)) and the
same output files (`packedSimPath/sim_%04d/source_%04d.uni`, GAN/multipassGAN-out.py:616).  Differences:
  * weights: like the reference (:153-186,367-386) `load_model_test_N` / `load_model_no_N` restore
    `basePath/test_%04d/model_%04d.ckpt` (`loadEmas 1`: `model_ema_%04d.ckpt`), read by `tfckpt.py` without
    TensorFlow (checkpoint V2 index + data files).  Two extensions for runs without a trained model:
    `randomInit <seed>` selects deterministic random-init weights and `weightsNpz <file>` loads a
    {variable name: array} archive with the reference's variable names.
  * PNG previews (scipy.misc.imsave, gone from scipy) are not written.
"""
import os
import sys
import time

import numpy as np


class Params:
    """tools_wscale/paramhelpers.py:16-37 semantics on an explicit argv."""

    def __init__(self, argv):
        self.argv = list(argv)
        self.used = [0] * len(self.argv)
        self.values = {}

    def get(self, name, default):
        v = default
        for i in range(1, len(self.argv)):
            if self.argv[i].lower() == name.lower() and i + 1 < len(self.argv):
                self.used[i] = self.used[i + 1] = 1
                v = self.argv[i + 1]
        self.values[name] = v
        return v

    def check_unused(self):
        bad = [(i, a) for i, a in enumerate(self.argv) if i >= 1 and not self.used[i]]
        for i, a in bad:
            print("Error: param %d '%s' not used!" % (i, a))
        if bad:
            raise SystemExit(1)


def load_frames(sim_path, frame_min, frame_max, use_velocities, vel_scale):
    """FluidDataLoader with multi_file_list ['density','velocity'] (GAN/multipassGAN-out.py:116-138): frames
    [Z,Y,X,C], C = (d, vx, vy, vz); velocity channels scaled by velScale (:138)."""
    from . import uni
    frames, head0 = [], None
    for f in range(frame_min, frame_max):
        head, dens = uni.read_uni(os.path.join(sim_path, "density_low_%04d.uni" % f))
        head0 = head0 or head
        chans = [dens.astype(np.float32)]
        if use_velocities:
            _, vel = uni.read_uni(os.path.join(sim_path, "velocity_low_%04d.uni" % f))
            chans.append(vel.astype(np.float32) * np.float32(vel_scale))
        frames.append(np.concatenate(chans, axis=-1))
    return frames, head0


def main(argv=None):
    argv = sys.argv if argv is None else argv
    ph = Params(argv)
    g = ph.get
    out_flag = int(g("out", 1))
    basePath = g("basePath", "../2ddata_gan/")
    randSeed = int(g("randSeed", 1))
    load = {i: (int(g("load_model_test_%d" % i, -1)), int(g("load_model_no_%d" % i, -1))) for i in (1, 2, 3)}
    simSizeLow = int(g("simSize", 64))
    tileSizeLow = int(g("tileSize", 16))
    upRes = int(g("upRes", 4))
    packedSimPath = g("packedSimPath", "../2ddata_sim/")
    fromSim = int(g("fromSim", 1000))
    frame_min = int(g("frame_min", 0))
    for name in ("genModel", "discModel", "testPathStartNo", "change_velocity", "upsamplingMode", "upsampledData",
                 "useVorticities", "useFlags", "useK_Eps_Turb", "usePixelShuffle", "use_mb_stddev"):
        g(name, 0)  # accepted for compatibility; they do not change the shipped apply path
    # GAN/multipassGAN-out.py:96-97 exports CUDA_VISIBLE_DEVICES = gpu before TensorFlow starts; CUDA may already be
    # initialised in this process, so the flag selects the device ordinal instead
    gpu = int(str(g("gpu", "0")).split(",")[0])
    load_emas = int(g("loadEmas", 0)) != 0  # model_ema_%04d.ckpt instead of model_%04d.ckpt (:156-160)
    batch_norm = int(g("batchNorm", 0)) != 0
    pixel_norm = int(g("pixelNorm", 1)) != 0
    useVelocities = int(g("useVelocities", 0))
    transposeAxis = int(g("transposeAxis", 0))
    frame_max = int(g("frame_max", 200))
    genUni = int(g("genUni", 0))
    addBicubic = int(g("addBicubicUpsample", 0)) != 0
    firstNNArch = int(g("firstNNArch", 1)) != 0
    upsampleMode = int(g("upsampleMode", 1))
    velScale = float(g("velScale", 1.0))
    specs = {}
    from . import pipeline as P
    for i in (1, 2, 3):
        specs[i] = P.NetSpec(use_res_net=int(g("use_res_net%d" % i, 0)) != 0, add_adj_idcs=int(g("add_adj_idcs%d" % i, 0)) != 0,
                             startFms=int(g("startFms%d" % i, 512)), maxFms=int(g("maxFms%d" % i, 256)),
                             filterSize=int(g("filterSize%d" % i, 3)), first_nn_arch=(firstNNArch and i == 1))
    random_init = g("randomInit", None)        # extension, see module docstring
    weights_npz = g("weightsNpz", None)        # extension
    precision = g("precision", "fp16")         # extension: fp16 | bf16 | fp32
    io_threads = int(g("ioThreads", 4))        # extension: gzip writer threads of the frame pipeline (io_pipeline.py)
    range_check = int(g("rangeCheck", 0)) != 0  # extension: validation mode, abort on saturated 16-bit activations
    uni_chunk_mb = int(g("uniChunkMB", 0))     # extension: >0 = multi-member gzip output deflated on ioThreads threads
    ph.check_unused()
    if tileSizeLow != simSizeLow:
        raise SystemExit("multipassGAN-out: the apply path slices whole frames (tileSize must equal simSize, as in "
                         "GAN/example_run_output.py)")
    nets = tuple(i for i in (1, 2, 3) if load[i][0] != -1)
    if not nets:
        print("At least one network has to be loaded.")
        raise SystemExit(1)
    if not useVelocities:
        raise SystemExit("multipassGAN-out: the shipped generators take (density, vx, vy, vz): useVelocities 1 is required")
    import torch
    if torch.cuda.is_available():
        if gpu >= torch.cuda.device_count():
            raise SystemExit("multipassGAN-out: gpu %d requested, %d visible" % (gpu, torch.cuda.device_count()))
        torch.cuda.set_device(gpu)
    if weights_npz:
        arch = np.load(weights_npz)
        weights = {i: {k: arch[k] for k in arch.files if k.startswith("gen_%d/" % i)} for i in nets}
    elif random_init is not None:
        weights = P.make_weights_out(simSizeLow, int(random_init), upRes=upRes, specs=specs, nets=nets,
                                     pixel_norm=pixel_norm, batch_norm=batch_norm, upsampleMode=upsampleMode,
                                     addBicubicUpsample=addBicubic)
    else:
        # the reference's own path (GAN/multipassGAN-out.py:153-186,367-386): basePath/test_%04d/model[_ema]_%04d.ckpt,
        # Saver keys = variable names without the `gen_N/` scope
        from . import tfckpt
        names = P.make_weights_out(simSizeLow, 0, upRes=upRes, specs=specs, nets=nets, pixel_norm=pixel_norm,
                                   batch_norm=batch_norm, upsampleMode=upsampleMode, addBicubicUpsample=addBicubic)
        weights = {}
        for i in nets:
            prefix = os.path.join(basePath, "test_%04d" % load[i][0],
                                  ("model_ema_%04d.ckpt" if load_emas else "model_%04d.ckpt") % load[i][1])
            if not os.path.exists(prefix + ".index"):
                raise SystemExit("multipassGAN-out: checkpoint %s.index not found; pass `randomInit <seed>` or "
                                 "`weightsNpz <file>` to run without a trained model" % prefix)
            weights[i] = tfckpt.load_generator_weights(prefix, sorted(names[i]), "gen_%d" % i)
            for n, ref in names[i].items():
                if weights[i][n].shape != ref.shape:
                    raise SystemExit("multipassGAN-out: %s: '%s' has shape %s, the graph built from the flags needs %s"
                                     % (prefix, n, weights[i][n].shape, ref.shape))
            print("Model %d restored from %s." % (i, prefix))
    sim_path = os.path.join(packedSimPath, "sim_%04d" % fromSim)
    mp = P.MultiPassOut(simSizeLow, weights, upRes=upRes, specs=specs, precision=precision, transposeAxis=transposeAxis,
                        threshold=P.THRESHOLD if genUni else 0.0, pixel_norm=pixel_norm, batch_norm=batch_norm,
                        upsampleMode=upsampleMode, addBicubicUpsample=addBicubic, device=gpu, range_check=range_check)
    if range_check:
        print("rangeCheck 1: every 16-bit layer output is scanned for saturated values (validation mode, slower)")
    S = simSizeLow * upRes
    print("*****OUTPUT ONLY*****")
    from . import io_pipeline, uni
    # header of frame 0 for every output, like the reference (GAN/multipassGAN-out.py:629, :606-610)
    head_path = os.path.join(sim_path, "density_low_%04d.uni" % 0)
    if not os.path.exists(head_path):
        head_path = os.path.join(sim_path, "density_low_%04d.uni" % frame_min)
    head, _ = uni.read_uni(head_path)
    head = dict(head)
    head["dimX"] = head["dimY"] = head["dimZ"] = S

    def load(f):
        return load_frames(sim_path, f, f + 1, useVelocities, velScale)[0][0]

    def compute(f, x):
        t0 = time.time()
        vol = mp(x)
        host = vol.cpu().numpy() if genUni else None  # the D2H copy synchronises
        torch.cuda.synchronize()
        print("%d  time for %d network(s): %.6f" % (f, len(nets), time.time() - t0))
        return host

    def store(f, host):
        if host is None:
            return
        out_path = os.path.join(sim_path, "source_%04d.uni" % f)
        if uni_chunk_mb > 0:
            io_pipeline.write_uni_parallel(out_path, head, host, threads=io_threads, chunk_bytes=uni_chunk_mb << 20)
        else:
            uni.write_uni(out_path, head, host)
        print("stored .uni file")

    stats = io_pipeline.FramePipeline(load, compute, store, prefetch=2, readers=1, writers=io_threads).run(
        range(frame_min, frame_max))
    print("frames %d: wall %.2f s (load %.2f, networks %.2f, store %.2f s of thread time)" % (
        stats["frames"], stats["wall_s"], stats["load_s"], stats["compute_s"], stats["store_s"]))
    return 0
