#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON.

Run in the authoring container only (needs /root/reference; the GPU box and the tests never do):

    python tests/golden/make_golden.py

Nothing from the reference is copied into the repo: the function bodies are read from
/root/reference at generation time (ast -> compile -> exec) and run on top of test doubles:

  pipeline_*.npz   generate3DUniForNewNetwork of GAN/multipassGAN-out.py:390-618 and
                   GAN/multipassGAN-4x.py:1090-1169 with `sess.run` replaced by the deterministic
                   stand-in row functions of tests/golden/standin.py.  PINS the zoom / slice-axis
                   permutation / velocity-channel swaps / adjacent-slice channels / slice batching /
                   inter-pass transposes / threshold, i.e. SURVEY §8 rows a14-a18, against the real code
                   (scipy.ndimage.zoom is the real library call).
  net_*.npz        tools_wscale/GAN.py (the real class) + resBlock / growBlockGen / growing_gen /
                   gen_resnet / disc_binclass of the scripts, executed eagerly on the numpy TF1 shim
                   (tests/golden/tf1_numpy_shim.py).  PINS layer wiring, variable names and shapes,
                   wscale constants, bias/BN/activation order and the cursor quirks (App. D) against the
                   real code; the op arithmetic itself stays a restatement of TF semantics (TensorFlow
                   cannot be installed here) - see the shim's docstring.
"""
import ast
import importlib.util
import json
import math
import os
import sys
import types
import zlib

import numpy as np
import scipy.ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import standin  # noqa: E402
import tf1_numpy_shim as tfs  # noqa: E402
import mpgan_b200  # noqa: E402,F401
from mpgan_b200 import synth, weights as W  # noqa: E402


def ref_functions(path, names):
    """Source of the named top-level functions of a reference script, compiled on its own."""
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in picked} == set(names), (path, names)
    mod = ast.Module(body=picked, type_ignores=[])
    return compile(mod, path, "exec")


# =============================================================================== pipelines
class _StubSess:
    def __init__(self, fn_by_sampler, log):
        self.fn, self.log = fn_by_sampler, log

    def run(self, sampler, feed_dict=None, **kw):
        xs = np.array(feed_dict["x"])
        ys = np.array(feed_dict["y"]) if "y" in feed_dict else None
        self.log.append((sampler, xs, ys))
        return self.fn[sampler](xs, ys)


class _Uni:
    def __init__(self):
        self.written = {}

    def readUni(self, path):
        return {"dimX": 0, "dimY": 0, "dimZ": 0}, None

    def writeUni(self, path, head, data):
        self.written[os.path.basename(path)] = (dict(head), np.array(data))


def run_out_pipeline(x_vol, u, nets, transposeAxis, add_adj1):
    """GAN/multipassGAN-out.py:390-618 on x_vol [L,L,L,4] with stand-in networks `nets` (subset of 1,2,3)."""
    L = x_vol.shape[0]
    S = L * u
    code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-out.py"), ["generate3DUniForNewNetwork"])
    log, uni = [], _Uni()
    fns = {
        "sampler": lambda xs, ys: standin.net_first(xs, L, u, 6 if add_adj1 else 4),
        "sampler_2": lambda xs, ys: standin.net_refine(xs, ys, L, u, 2),
        "sampler_3": lambda xs, ys: standin.net_refine(xs, ys, L, u, 3),
    }
    ns = dict(
        np=np, scipy=scipy, time=__import__("time"), tf=types.SimpleNamespace(RunMetadata=lambda: None),
        x_3d=np.array(x_vol, dtype=np.float32)[None], upRes=u, simSizeLow=L, simSizeHigh=S, tileSizeHigh=S,
        n_inputChannels=4, n_input=L * L * 4, n_output=S * S, transposeAxis=transposeAxis,
        add_adj_idcs1=add_adj1, add_adj_idcs2=False, add_adj_idcs3=False,
        load_model_test_1=0 if 1 in nets else -1, load_model_test_2=0 if 2 in nets else -1,
        load_model_test_3=0 if 3 in nets else -1,
        load_model_no_1=0 if 1 in nets else -1, load_model_no_2=0 if 2 in nets else -1,
        load_model_no_3=0 if 3 in nets else -1,
        sess=_StubSess(fns, log), sampler="sampler", sampler_2="sampler_2", sampler_3="sampler_3",
        x="x", y="y", percentage="percentage", train="train",
        save_img_3d=lambda *a, **k: None, save_img=lambda *a, **k: None,
        uniio=uni, packedSimPath="/nonexistent/", fromSim=1000, frame_min=0, generateUni=True,
    )
    exec(code, ns)
    ns["generate3DUniForNewNetwork"](imageindex=0, outPath="/nonexistent/", head={"dimX": 0, "dimY": 0, "dimZ": 0})
    (head, vol), = uni.written.values()
    assert head["dimX"] == S and vol.shape == (S, S, S)
    feeds = {}
    for name in ("sampler", "sampler_2", "sampler_3"):
        xs = [a for (s, a, b) in log if s == name]
        if xs:
            feeds[name + "_x"] = np.concatenate(xs, axis=0)
    return vol.astype(np.float32), feeds


def run_4x_pass(mode, u, x_3d, x_2=None):
    """GAN/multipassGAN-4x.py:1090-1169, one frame, upsampleFirst=1, genUni=1."""
    L = x_3d.shape[0]
    S = L * u
    code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-4x.py"), ["generate3DUniForNewNetwork"])
    log, uni = [], _Uni()
    fn = (lambda xs, ys: standin.net_first(xs, L, u, 4)) if mode == 2 else (lambda xs, ys: standin.net_fullres(xs, S))
    ns = dict(
        np=np, scipy=scipy, time=__import__("time"), upsampling_mode=mode, upsampleFirst=1, upRes=u,
        x_3d=np.array(x_3d, dtype=np.float32)[None], x_2=None if x_2 is None else np.array(x_2, np.float32)[None],
        simSizeLow=L, simSizeHigh=S, n_inputChannels=4, n_input=(L * L if mode == 2 else S * S) * 4,
        sess=_StubSess({"sampler": fn}, log), sampler="sampler", x="x", keep_prob="keep_prob", dropoutOutput=1.0,
        train="train", save_img_3d=lambda *a, **k: None, uniio=uni, packedSimPath="/nonexistent/", fromSim=1000,
        frame_min=0, generateUni=True, test_path="/nonexistent/",
    )
    exec(code, ns)
    ns["generate3DUniForNewNetwork"](imageindex=0, outPath="/nonexistent/")
    (head, vol), = uni.written.values()
    feeds = np.concatenate([a for (s, a, b) in log], axis=0)
    return vol.astype(np.float32), feeds


def make_pipeline_fixtures():
    out = {}
    L, u = 4, 4  # S = 16: 16 slices -> two batches of 8 (net 1) / eight batches of 2 (nets 2, 3)
    x = synth.synthetic_volume(L, seed=21)
    for ta, nets, adj in ((0, (1, 2), True), (0, (1, 2, 3), True), (1, (1, 2, 3), False), (3, (1, 2), True),
                          (2, (1,), True), (0, (1,), False)):
        vol, feeds = run_out_pipeline(x, u, nets, ta, adj)
        key = "out_ta%d_n%s_adj%d" % (ta, "".join(map(str, nets)), int(adj))
        out[key + "_vol"] = vol
        for k, v in feeds.items():
            out[key + "_" + k] = v.astype(np.float32)
    # 4x: pass 1 (mode 2) -> thresholded volume -> pass 2 (mode 1) and the mode-3 variant
    p1, f1 = run_4x_pass(2, u, x)
    vel = x[..., 1:4] * u  # GAN/multipassGAN-4x.py:277-278 (velScale 1.0)
    p2, f2 = run_4x_pass(1, u, vel, x_2=p1[..., None])
    p3, f3 = run_4x_pass(3, u, vel, x_2=p1[..., None])
    out.update({"x": x, "L": L, "u": u, "x4_p1_vol": p1, "x4_p1_feed": f1, "x4_p2_vol": p2, "x4_p2_feed": f2,
                "x4_p3_vol": p3, "x4_p3_feed": f3})
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)
    print("pipeline.npz:", sorted(out))


# =============================================================================== networks
def load_ref_gan():
    """Import the REAL tools_wscale/GAN.py with `tensorflow` / `keras` resolved to the numpy shim."""
    sys.modules["tensorflow"] = tfs
    keras = types.ModuleType("keras")
    keras.backend = tfs.keras_backend
    sys.modules["keras"] = keras
    spec = importlib.util.spec_from_file_location("ref_GAN", os.path.join(REF, "tools_wscale", "GAN.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Provider(dict):
    """Variable values keyed by the reference's full variable name, generated on first request with the
    repo's deterministic initialiser (weights.py) + non-trivial BN statistics."""

    def __init__(self, seed):
        super().__init__()
        self.seed = seed
        self.kinds = {}

    def __contains__(self, name):
        return True

    def __missing__(self, name):
        raise KeyError(name)


def _provide(seed):
    store = {}

    def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True):
        full = "/".join(tfs.STATE.scopes + [name])
        shape = tuple(int(s) for s in shape)
        tfs.STATE.requested[full] = shape
        if full not in store:
            if name in ("gamma", "beta", "moving_mean", "moving_variance"):
                base = {"gamma": 1.0, "beta": 0.0, "moving_mean": 0.0, "moving_variance": 1.0}[name]
                v = W.init_variable(seed, full, shape, ("const", base))
                v = W.randomize_bn_stats({full: v}, seed)[full]
            elif initializer[0] == "normal":
                v = W.init_variable(seed, full, shape, "normal")
            elif initializer[0] == "const":
                v = W.init_variable(seed, full, shape, ("const", initializer[1]))
            else:
                raise NotImplementedError(initializer)
            store[full] = v
        assert store[full].shape == shape
        return tfs.T(store[full])

    return store, get_variable


def _names_blob(requested):
    return json.dumps([[k, list(v)] for k, v in requested.items()])


def _wsum(store):
    return float(sum(np.float64(zlib.crc32(np.ascontiguousarray(v).tobytes())) for v in store.values()))


def make_net_fixtures():
    ref_gan = load_ref_gan()
    out_code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-out.py"), ["resBlock", "growBlockGen", "growing_gen"])
    x4_code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-4x.py"), ["resBlock", "gen_resnet", "disc_binclass"])
    fixtures = {}
    rng = np.random.default_rng(5)

    def run_out(tag, seed, L, u, firstGen, spec, pixel_norm=True, addBicubic=True, use_bn=False):
        S = L * u
        store, getv = _provide(seed)
        tfs.reset({})
        tfs.get_variable = getv
        C = 4
        ns = dict(tf=tfs, GAN=ref_gan.GAN, lrelu=ref_gan.lrelu, np=np, math=math, train=False, pixel_norm=pixel_norm,
                  usePixelShuffle=False, upsampleMode=1, addBicubicUpsample=addBicubic, n_inputChannels=C,
                  tileSizeLow=L, tileSizeHigh=S, n_output=S * S, rbId=0, print=lambda *a, **k: None)
        exec(out_code, ns)
        B = 2
        cin = C + (2 if spec["add_adj_idcs"] else 0)
        idx = spec["idx"]
        x_rows = rng.random((B, L * L * (cin if firstGen else C)), dtype=np.float32)
        y_rows = None
        with tfs.variable_scope("gen_%d" % idx):
            if firstGen:
                _in = tfs.T(x_rows)
            else:
                # sampler wiring of GAN/multipassGAN-out.py:357 (module-level code, restated here)
                y_rows = rng.random((B, S * S), dtype=np.float32)
                lo = tfs.reshape(tfs.T(x_rows), [-1, L, L, C])
                _in = tfs.concat((tfs.reshape(tfs.T(y_rows), [-1, S, S, 1]),
                                  tfs.image.resize_images(lo, tfs.constant([S, S]), method=1)), axis=3)
            res = ns["growing_gen"](_in, percentage=None, use_batch_norm=use_bn, reuse=tfs.AUTO_REUSE,
                                    currentUpres=int(round(math.log(u, 2))), train=False, output=True,
                                    firstGen=firstGen, filterSize=spec["filterSize"], startFms=spec["startFms"],
                                    maxFms=spec["maxFms"], add_adj_idcs=spec["add_adj_idcs"],
                                    first_nn_arch=spec["first_nn_arch"], use_res_net=spec["use_res_net"])
        fixtures[tag + "_x"] = x_rows
        if y_rows is not None:
            fixtures[tag + "_y"] = y_rows
        fixtures[tag + "_out"] = res.a
        fixtures[tag + "_vars"] = _names_blob(tfs.STATE.requested)
        fixtures[tag + "_wsum"] = _wsum(store)
        fixtures[tag + "_cfg"] = json.dumps(dict(seed=seed, L=L, u=u, firstGen=firstGen, spec=spec, pixel_norm=pixel_norm,
                                                 addBicubic=addBicubic))
        print(tag, res.a.shape, "vars", len(tfs.STATE.requested), "|out| max", float(np.abs(res.a).max()))

    run_out("out_net1", 31, 8, 4, True, dict(idx=1, use_res_net=True, add_adj_idcs=True, startFms=32, maxFms=32,
                                             filterSize=3, first_nn_arch=True))
    run_out("out_net1_u8", 32, 4, 8, True, dict(idx=1, use_res_net=True, add_adj_idcs=True, startFms=128, maxFms=64,
                                                filterSize=3, first_nn_arch=True))
    run_out("out_net2", 33, 4, 4, False, dict(idx=2, use_res_net=True, add_adj_idcs=False, startFms=64, maxFms=64,
                                              filterSize=5, first_nn_arch=False))
    run_out("out_net3", 34, 4, 4, False, dict(idx=3, use_res_net=False, add_adj_idcs=False, startFms=64, maxFms=32,
                                              filterSize=5, first_nn_arch=False))

    def run_4x(tag, seed, L, u, mode, use_bn=True):
        S = L * u
        store, getv = _provide(seed)
        tfs.reset({})
        tfs.get_variable = getv
        ns = dict(tf=tfs, GAN=ref_gan.GAN, lrelu=ref_gan.lrelu, np=np, dataDimension=2, train=False, rbId=0,
                  upsampling_mode=mode, tileSizeLow=L, tileSizeHigh=S, n_inputChannels=4, n_output=S * S,
                  useAvgDepool=False, upRes=u, n_input=(L * L if mode == 2 else S * S) * 4, bn_decay=0.999,
                  print=lambda *a, **k: None)
        exec(x4_code, ns)
        B = 2
        n_in = ns["n_input"]
        x_rows = (rng.random((B, n_in), dtype=np.float32) * np.tile([1, .5, .5, .5], n_in // 4)).astype(np.float32)
        res = ns["gen_resnet"](tfs.T(x_rows), reuse=False, use_batch_norm=use_bn, train=False)
        fixtures[tag + "_x"] = x_rows
        fixtures[tag + "_out"] = res.a
        fixtures[tag + "_vars"] = _names_blob(tfs.STATE.requested)
        fixtures[tag + "_wsum"] = _wsum(store)
        fixtures[tag + "_cfg"] = json.dumps(dict(seed=seed, L=L, u=u, mode=mode, batch_norm=use_bn))
        print(tag, res.a.shape, "vars", len(tfs.STATE.requested), "|out| max", float(np.abs(res.a).max()))
        if mode == 2:
            # spatial discriminator on (x, G(x)) -- inference-mode BN (moving statistics)
            tfs.STATE.requested = {}
            g_rows = np.maximum(res.a, 0).astype(np.float32)
            d = ns["disc_binclass"](tfs.T(x_rows), tfs.T(g_rows), reuse=False, use_batch_norm=use_bn, train=False)
            fixtures[tag + "_disc_y"] = g_rows
            fixtures[tag + "_disc_logits"] = d[0].a
            for i in range(1, 5):
                fixtures[tag + "_disc_d%d" % i] = d[i].a
            fixtures[tag + "_disc_vars"] = _names_blob(tfs.STATE.requested)
            fixtures[tag + "_disc_wsum"] = _wsum(store)
            print(tag + "_disc", d[0].a.ravel(), [t.a.shape for t in d[1:]])

    run_4x("x4_mode2", 41, 16, 4, 2)   # tile 16 -> 64: the training-tile shape, so the disc head is 8x8x256
    run_4x("x4_mode1", 42, 4, 4, 1)
    run_4x("x4_mode2_nobn", 43, 4, 4, 2, use_bn=False)
    np.savez_compressed(os.path.join(HERE, "nets.npz"), **fixtures)
    print("nets.npz written:", len(fixtures), "arrays")


# =============================================================================== 8x trainer: growing discriminator
def make_growdisc_fixtures():
    """growing_disc / growBlockDisc / lerp of GAN/multipassGAN-8x.py (:596-597, 752-866) executed on the numpy TF1 shim with
    the module-level flags of the shipped first-network training command (GAN/example_run_training.py:4) at small sizes."""
    ref_gan = load_ref_gan()
    code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-8x.py"), ["lerp", "growBlockDisc", "growing_disc"])
    fixtures = {}
    rng = np.random.default_rng(8)

    def run(tag, seed, L, u, C, start_fms, max_fms, first_nn_arch, percentages, filterSize=3, upsampling_mode=2):
        S = L * u
        store, getv = _provide(seed)
        tfs.reset({})
        tfs.get_variable = getv
        ns = dict(tf=tfs, GAN=ref_gan.GAN, lrelu=ref_gan.lrelu, np=np, math=math, tileSizeLow=L, tileSizeHigh=S, upRes=u,
                  n_inputChannels=C, upsampling_mode=upsampling_mode, upsampleMode=1, start_fms=start_fms, max_fms=max_fms,
                  filterSize=filterSize, first_nn_arch=first_nn_arch, useVelInTDisc=False, bn_decay=0.999,
                  use_mb_stddev=False, gn=lambda x, gstr: x, print=lambda *a, **k: None)
        exec(code, ns)
        B = 2
        x_rows = rng.random((B, L * L * C), dtype=np.float32)
        y_rows = rng.random((B, S * S), dtype=np.float32)
        fixtures[tag + "_x"], fixtures[tag + "_y"] = x_rows, y_rows
        for k, pct in enumerate(percentages):
            tfs.STATE.requested = {}
            logits, feats = ns["growing_disc"](tfs.T(y_rows), tfs.T(x_rows), tfs.T(np.float32(pct)), reuse=tfs.AUTO_REUSE,
                                               use_batch_norm=False, train=False, currentUpres=int(round(math.log(u, 2))))
            fixtures["%s_p%d_logits" % (tag, k)] = logits.a
            for i, f in enumerate(feats):  # feature layers: full tensors for one percentage, (sum, sum |.|) for the others
                if k == 1:
                    fixtures["%s_p%d_feat%d" % (tag, k, i)] = f.a
                else:
                    fixtures["%s_p%d_feat%d_sums" % (tag, k, i)] = np.array([f.a.sum(dtype=np.float64), np.abs(f.a).sum(dtype=np.float64)])
            print(tag, pct, logits.a.ravel(), len(feats), "feature layers")
        fixtures[tag + "_vars"] = _names_blob(tfs.STATE.requested)
        fixtures[tag + "_wsum"] = _wsum(store)
        fixtures[tag + "_cfg"] = json.dumps(dict(seed=seed, L=L, u=u, C=C, start_fms=start_fms, max_fms=max_fms,
                                                 first_nn_arch=first_nn_arch, percentages=list(percentages),
                                                 filterSize=filterSize, upsampling_mode=upsampling_mode))

    tcode = ref_functions(os.path.join(REF, "GAN", "multipassGAN-8x.py"), ["lerp", "growBlockDisc", "growing_disc_tempo"])

    def run_tempo(tag, seed, L, u, start_fms, max_fms, first_nn_arch, percentages, filterSize=3, upsampling_mode=2):
        """growing_disc_tempo (:868-923): the unconditional critic of three aligned frames, input [B, S*S, 3]."""
        S = L * u
        store, getv = _provide(seed)
        tfs.reset({})
        tfs.get_variable = getv
        ns = dict(tf=tfs, GAN=ref_gan.GAN, lrelu=ref_gan.lrelu, np=np, math=math, tileSizeLow=L, tileSizeHigh=S, upRes=u,
                  upsampling_mode=upsampling_mode, upsampleMode=1, start_fms=start_fms, max_fms=max_fms, filterSize=filterSize,
                  first_nn_arch=first_nn_arch, useVelInTDisc=False, bn_decay=0.999, use_mb_stddev=False, gn=lambda x, gstr: x,
                  print=lambda *a, **k: None)
        exec(tcode, ns)
        frames = rng.random((2, S * S, 3), dtype=np.float32)
        fixtures[tag + "_frames"] = frames
        for k, pct in enumerate(percentages):
            tfs.STATE.requested = {}
            logits = ns["growing_disc_tempo"](tfs.T(frames), tfs.T(np.float32(pct)), reuse=tfs.AUTO_REUSE, use_batch_norm=False,
                                              train=False, currentUpres=int(round(math.log(u, 2))))
            fixtures["%s_p%d_logits" % (tag, k)] = logits.a
            print(tag, pct, logits.a.ravel())
        fixtures[tag + "_vars"] = _names_blob(tfs.STATE.requested)
        fixtures[tag + "_wsum"] = _wsum(store)
        fixtures[tag + "_cfg"] = json.dumps(dict(seed=seed, L=L, u=u, C=0, start_fms=start_fms, max_fms=max_fms,
                                                 first_nn_arch=first_nn_arch, percentages=list(percentages),
                                                 filterSize=filterSize, upsampling_mode=upsampling_mode))

    gen_code = ref_functions(os.path.join(REF, "GAN", "multipassGAN-8x.py"), ["lerp", "resBlock", "growBlockGen", "growing_gen"])

    def run_gen(tag, seed, L, u, C, start_fms, max_fms, percentages, first_nn_arch=True, upsampling_mode=2, filterSize=3):
        """growing_gen in TRAINING mode (output=False: per-stage density outputs blended with lerp, :700-750)."""
        S = L * u
        store, getv = _provide(seed)
        tfs.reset({})
        tfs.get_variable = getv
        ns = dict(tf=tfs, GAN=ref_gan.GAN, lrelu=ref_gan.lrelu, np=np, math=math, tileSizeLow=L, tileSizeHigh=S, upRes=u,
                  n_inputChannels=C, n_output=S * S, upsampling_mode=upsampling_mode, upsampleMode=1, start_fms=start_fms,
                  max_fms=max_fms, filterSize=filterSize, first_nn_arch=first_nn_arch, use_res_net=True, pixel_norm=True,
                  usePixelShuffle=False,
                  addBicubicUpsample=True, dataDimension=2, bn_decay=0.999, train=False, rbId=0, print=lambda *a, **k: None)
        exec(gen_code, ns)
        # upsampling_mode 1 / 3: the network input is the high-res concat(first-pass density, resized low-res fields) :684
        x_rows = rng.random((2, L * L * C) if upsampling_mode == 2 else (2, S * S * (C + 1)), dtype=np.float32)
        fixtures[tag + "_x"] = x_rows
        for k, pct in enumerate(percentages):
            tfs.STATE.requested = {}
            res = ns["growing_gen"](tfs.T(x_rows), tfs.T(np.float32(pct)), reuse=tfs.AUTO_REUSE, use_batch_norm=False, train=False,
                                    currentUpres=int(round(math.log(u, 2))), output=False)
            fixtures["%s_p%d_out" % (tag, k)] = res.a
            print(tag, pct, res.a.shape, float(np.abs(res.a).max()))
        fixtures[tag + "_vars"] = _names_blob(tfs.STATE.requested)
        fixtures[tag + "_wsum"] = _wsum(store)
        fixtures[tag + "_cfg"] = json.dumps(dict(seed=seed, L=L, u=u, C=C, start_fms=start_fms, max_fms=max_fms,
                                                 percentages=list(percentages), first_nn_arch=first_nn_arch,
                                                 upsampling_mode=upsampling_mode, filterSize=filterSize))

    run_gen("gg_first", 83, 4, 8, 6, 32, 32, (0.4, 1.3, 2.75, 3.0))
    run("gd_first", 81, 4, 8, 6, 32, 32, True, (0.4, 1.3, 2.75, 3.0))
    run("gd_plain", 82, 4, 4, 4, 32, 16, False, (0.5, 1.6, 2.0))
    # the second (refinement) network's training graph: upsampling_mode 1, firstNNArch 0 (GAN/example_run_training.py:7)
    run_gen("gg_second", 84, 2, 8, 4, 32, 32, (0.4, 1.3, 2.75, 3.0), first_nn_arch=False, upsampling_mode=1, filterSize=5)
    run("gd_second", 85, 2, 8, 4, 32, 32, False, (0.4, 1.3, 2.75, 3.0), filterSize=5, upsampling_mode=1)
    # the temporal critic of both shipped training commands (lambda_t 1.0): first network / refinement network
    run_tempo("gt_first", 86, 4, 8, 32, 32, True, (0.4, 1.3, 2.75, 3.0))
    run_tempo("gt_second", 87, 2, 8, 32, 32, False, (0.4, 1.3, 2.75, 3.0), filterSize=5, upsampling_mode=1)
    # tensorResample (:545-594): the bilinear re-sampling of a frame at CPU-advected positions in front of the temporal critic
    rcode = ref_functions(os.path.join(REF, "GAN", "multipassGAN-8x.py"), ["tensorResample"])
    rns = dict(tf=tfs, np=np, pow=pow, int=int, bool=bool, range=range, len=len)
    exec(rcode, rns)
    Bn, Hn = 3, 8
    val = rng.random((Bn, Hn, Hn, 2), dtype=np.float32)
    base = np.stack(np.meshgrid(np.arange(Hn) + 0.5, np.arange(Hn) + 0.5, indexing="ij"), axis=-1)[None]
    posn = (base + rng.normal(0.0, 1.2, (Bn, Hn, Hn, 2))).astype(np.float32)      # some positions leave the tile
    fixtures["resample_value"], fixtures["resample_pos"] = val, posn
    fixtures["resample_out"] = rns["tensorResample"](tfs.T(val), tfs.T(posn)).a
    ident = rns["tensorResample"](tfs.T(val), tfs.T(np.broadcast_to(base, (Bn, Hn, Hn, 2)).copy())).a
    assert np.abs(ident - val).max() < 1e-12            # cell centres reproduce the values
    print("resample:", fixtures["resample_out"].shape, float(np.abs(fixtures["resample_out"]).max()))
    np.savez_compressed(os.path.join(HERE, "growdisc.npz"), **fixtures)
    print("growdisc.npz written:", len(fixtures), "arrays")


def ref_methods(path, cls, names):
    """The named methods of a reference class, compiled as plain functions."""
    with open(path) as fh:
        tree = ast.parse(fh.read())
    c = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0]
    picked = [n for n in c.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in picked} == set(names)
    return compile(ast.Module(body=picked, type_ignores=[]), path, "exec")


def make_tile_fixtures():
    """TileCreator.createTiles / cutTile / concatTiles (tools_wscale/tilecreator_t.py; the module itself cannot be
    imported on py3.12 - `import imp` - so the three methods are compiled on their own and bound to a stub self)."""
    code = ref_methods(os.path.join(REF, "tools_wscale", "tilecreator_t.py"), "TileCreator",
                       ["createTiles", "cutTile", "concatTiles"])
    ns = dict(np=np)
    exec(code, ns)

    class Stub:
        def __init__(self, padding):
            self.padding = padding

        def TCError(self, msg):
            raise RuntimeError(msg)

    for name in ("createTiles", "cutTile", "concatTiles"):
        setattr(Stub, name, ns[name])
    rng = np.random.default_rng(77)
    frame = rng.random((1, 24, 40, 3), dtype=np.float32)
    out = {"frame": frame}
    # regular grid and overlapping grid (stride < tile).  padding > 0 is NOT pinned: the reference's
    # np.pad(currTile, [p,p,p,0], 'edge') (tilecreator_t.py:430) raises for a 4-D tile in numpy, i.e. the branch never ran
    for tag, (tile, stride, pad) in {"reg": ([1, 8, 8], -1, 0), "ovl": ([1, 12, 16], 4, 0), "ovl2": ([1, 8, 10], 6, 0)}.items():
        t = Stub(pad).createTiles(frame, list(tile), stride)
        out["tiles_" + tag] = np.asarray(t, np.float32)
    # stitch: 2 x 3 tiles of 12 x 16 with a 2-pixel border cropped (and the uncropped variant)
    tiles = rng.random((6, 1, 12, 16, 3), dtype=np.float32)
    out["stitch_in"] = tiles
    out["stitch_b2"] = np.asarray(Stub(0).concatTiles(tiles, [1, 2, 3], [0, 2, 2, 0]), np.float32)
    out["stitch_b0"] = np.asarray(Stub(0).concatTiles(tiles, [1, 2, 3], [0, 0, 0, 0]), np.float32)
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), **out)
    print("tiles.npz:", {k: v.shape for k, v in out.items()})


def make_uni_fixtures():
    """Two small grids written by the reference's own tools_wscale/uniio.writeUni (importable as-is)."""
    spec = importlib.util.spec_from_file_location("ref_uniio", os.path.join(REF, "tools_wscale", "uniio.py"))
    uniio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(uniio)
    rng = np.random.default_rng(3)
    dens = rng.random((3, 4, 5, 1), dtype=np.float32)
    vel = rng.standard_normal((3, 4, 5, 3)).astype(np.float32)
    from collections import OrderedDict
    def head(et):
        return OrderedDict([("dimX", 5), ("dimY", 4), ("dimZ", 3), ("gridType", 1 if et == 1 else 4), ("elementType", et),
                            ("bytesPerElement", 12 if et == 2 else 4), ("info", b"mantaflow test grid".ljust(252, b"\0")),
                            ("dimT", 0), ("timestamp", 1234567890123)])
    uniio.writeUni(os.path.join(HERE, "ref_density.uni"), head(1), dens)
    uniio.writeUni(os.path.join(HERE, "ref_velocity.uni"), head(2), vel)
    h, d = uniio.readUni(os.path.join(HERE, "ref_density.uni"))
    assert np.array_equal(d, dens)
    np.savez_compressed(os.path.join(HERE, "uni.npz"), dens=dens, vel=vel)
    print("ref_density.uni / ref_velocity.uni / uni.npz written")


def make_sampler_fixtures():
    """TileCreator.selectRandomTiles without augmentation (the default of GAN/multipassGAN-4x.py: dataAugmentation 0):
    the reference's own selectRandomTiles / getRandomDatum / getDatum / getRandomTile / cutTile / hasMinDensity /
    getTileDensity / splitSets, compiled from tools_wscale/tilecreator_t.py and bound to a stub that carries the state
    __init__ / addData would have set up for 2-D data (dim=2, dim_t=1). `randrange` is Python's random.randrange; the
    reference passes numpy floats to it (np.floor(bounds)), which Python >= 3.12 rejects, so the shim converts the
    (integral) bounds with int() -- the documented behaviour of the interpreters the reference ran on."""
    import random
    names = ["selectRandomTiles", "getRandomDatum", "getDatum", "getRandomTile", "cutTile", "hasMinDensity",
             "getTileDensity", "splitSets"]
    code = ref_methods(os.path.join(REF, "tools_wscale", "tilecreator_t.py"), "TileCreator", names)
    rnd = random.Random()

    def randrange(a, b):
        return rnd.randrange(int(a), int(b))

    ns = dict(np=np, randrange=randrange, DATA_KEY_LOW=0, DATA_KEY_HIGH=1, print=lambda *a, **k: None)
    exec(code, ns)

    class Stub:
        def TCError(self, msg):
            raise RuntimeError(msg)

    for name in names:
        setattr(Stub, name, ns[name])
    out = {}
    for tag, (T, L, u, nframes, dmin, part_train, seed) in {"a": (8, 24, 4, 10, 0.02, 0.9, 11), "b": (6, 16, 2, 5, 0.005, 0.6, 5)}.items():
        rng = np.random.default_rng(seed)
        S = L * u
        low = rng.random((nframes, 1, L, L, 4), dtype=np.float32)
        low[..., 0] *= (rng.random((nframes, 1, L, L)) < 0.15)       # sparse density: many tiles fail densityMinimum
        low[:, :, : L // 2, :, 0] = 0.0                               # an empty half so that retries really happen
        high = rng.random((nframes, 1, S, S, 1), dtype=np.float32)
        st = Stub()
        st.dim, st.dim_t, st.upres, st.premadeTiles, st.useDataAug = 2, 1, u, False, False
        st.densityMinimum = dmin
        st.tile_shape_low = np.array([1, T, T, 4])
        st.tile_shape_high = np.array([1, T * u, T * u, 1])
        st.data_flags = {0: dict(channels=4, isLabel=False), 1: dict(channels=1, isLabel=False)}
        st.data = {0: list(low), 1: list(high)}
        part_test = {0.9: 0.1, 0.6: 0.4}[part_train]  # TileCreator.__init__: part_x = partX / (partTrain + partTest + partVal)
        st.part_train, st.part_test = part_train / (part_train + part_test), part_test / (part_train + part_test)
        st.splitSets()
        rnd.seed(1000 + seed)
        xs, ys = [], []
        for call in range(3):
            bl, bh = st.selectRandomTiles(5, isTraining=True, augment=False)
            xs.append(np.asarray(bl, np.float32))
            ys.append(np.asarray(bh, np.float32))
        bl, bh = st.selectRandomTiles(4, isTraining=False, augment=False)
        out.update({tag + "_low": low, tag + "_high": high, tag + "_train_low": np.stack(xs), tag + "_train_high": np.stack(ys),
                    tag + "_test_low": np.asarray(bl, np.float32), tag + "_test_high": np.asarray(bh, np.float32),
                    tag + "_cfg": np.array([T, L, u, nframes, dmin, part_train, 1000 + seed], np.float64),
                    tag + "_borders": np.array(st.setBorders)})
    np.savez_compressed(os.path.join(HERE, "tilesampler.npz"), **out)
    print("tilesampler.npz:", {k: v.shape for k, v in out.items()})


def make_augment_fixtures():
    """TileCreator.selectRandomTiles(augment=True) -> generateTile (tools_wscale/tilecreator_t.py:491-546) with the data
    augmentation of the shipped 4x command line (GAN/example_run_output.py:6: `dataAugmentation 1 rot 1`, default
    minScale 0.85 / maxScale 1.15 / flip 1, GAN/multipassGAN-4x.py:95-102,263-264): random scaling (scipy.ndimage.zoom,
    order 1), a second random tile cut, a random 90-degree rotation and a random flip, with the velocity channels fixed up
    (scaleVelocities / rotate90Velocities / flipVelocities). The reference's own methods are compiled from the source
    and bound to a stub carrying the state __init__ / initDataAugmentation set up for 2-D data with channel layout
    'd,vx,vy,vz'. Two shims for today's interpreters, both documented behaviour of the versions the reference ran on:
    `randrange` gets int() bounds (numpy floats are rejected by Python >= 3.12) and `np.random.choice` of the ragged
    rotation list picks `lst[np.random.randint(0, len(lst))]` (numpy < 1.24 built a 1-D object array and drew one index)."""
    import random
    import types
    import scipy.ndimage
    names = ["selectRandomTiles", "generateTile", "getRandomDatum", "getDatum", "getRandomTile", "cutTile", "hasMinDensity",
             "getTileDensity", "splitSets", "special_aug", "scale", "scaleVelocities", "rotate90", "rotate90Velocities", "flip",
             "flipVelocities"]
    code = ref_methods(os.path.join(REF, "tools_wscale", "tilecreator_t.py"), "TileCreator", names)
    rnd = random.Random()

    def randrange(a, b):
        return rnd.randrange(int(a), int(b))

    class _Rand:  # np.random with the legacy ragged-list behaviour of choice()
        def __getattr__(self, k):
            return getattr(np.random, k)

        @staticmethod
        def choice(a):
            if isinstance(a, (int, np.integer)):
                return np.random.randint(0, a)
            return a[np.random.randint(0, len(a))]

    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    np_shim.random = _Rand()
    ns = dict(np=np_shim, scipy=scipy, randrange=randrange, DATA_KEY_LOW=0, DATA_KEY_HIGH=1, AOPS_KEY_ROTATE="rot",
              AOPS_KEY_SCALE="scale", AOPS_KEY_ROT90="rot90", AOPS_KEY_FLIP="flip", print=lambda *a, **k: None)
    exec(code, ns)

    class Stub:
        def TCError(self, msg):
            raise RuntimeError(msg)

    for name in names:
        setattr(Stub, name, ns[name])
    out = {}
    # "t3": THREE-frame tiles (TileCreator(dim_t=3), selectRandomTiles(..., tile_t=3) as selectRandomTempoTiles calls it,
    # :1391): the frames of a sequence are channel groups and special_aug applies the velocity fix-ups per frame
    cfgs = {"a": (8, 24, 4, 4, 0.02, 0.85, 1.15, 1, 1, 21), "b": (6, 20, 2, 5, 0.005, 1.0, 1.0, 1, 0, 22),
            "c": (8, 24, 2, 4, 0.02, 0.7, 1.3, 0, 1, 23), "t3": (6, 18, 2, 4, 0.02, 0.85, 1.15, 1, 1, 24)}
    for tag, (T, L, u, nframes, dmin, smin, smax, rot, flip, seed) in cfgs.items():
        dim_t = 3 if tag == "t3" else 1
        rng = np.random.default_rng(seed)
        S = L * u
        low = rng.random((nframes, 1, L, L, 4 * dim_t), dtype=np.float32)
        for k in range(dim_t):
            low[..., 4 * k + 1:4 * k + 4] -= 0.5
        low[..., 0] *= (rng.random((nframes, 1, L, L)) < 0.3)
        high = rng.random((nframes, 1, S, S, dim_t), dtype=np.float32)
        st = Stub()
        st.dim, st.dim_t, st.upres, st.premadeTiles, st.useDataAug = 2, dim_t, u, False, True
        st.densityMinimum = dmin
        st.tile_shape_low = np.array([1, T, T, 4])
        st.tile_shape_high = np.array([1, T * u, T * u, 1])
        vel = {"d": [0], "v": [[1, 2, 3]], "x": [], "o": [], "f": [], "k": [], "e": []}
        none = {"d": [0], "v": [], "x": [], "o": [], "f": [], "k": [], "e": []}
        st.c_lists = {0: vel, 1: none}
        st.data_flags = {0: dict(channels=4, isLabel=False, d=True, v=True, x=False), 1: dict(channels=1, isLabel=False, d=True, v=False, x=False)}
        st.aops = {k: {"rot": {}, "scale": {"v": st.scaleVelocities, "x": st.scaleVelocities},
                       "rot90": {"v": st.rotate90Velocities, "x": st.rotate90Velocities},
                       "flip": {"v": st.flipVelocities, "x": st.flipVelocities}} for k in (0, 1)}
        # initDataAugmentation(rot, minScale, maxScale, flip) :227-317
        st.do_rotation, st.do_rot90 = False, rot == 1
        z, nz = (2, 1), (1, 2)
        st.cube_rot = {2: [[], [z], [z, z], [nz]]}
        st.scaleFactor = [smin, smax]
        st.do_scaling = not (smin == 1 and smax == 1)
        st.do_flip = flip
        st.interpolation_order, st.fill_mode = 1, "constant"
        st.data = {0: list(low), 1: list(high)}
        st.part_train, st.part_test = 0.9, 0.1
        st.splitSets()
        rnd.seed(2000 + seed)
        np.random.seed(3000 + seed)
        xs, ys = [], []
        for call in range(3):
            bl, bh = st.selectRandomTiles(6, isTraining=True, augment=True, tile_t=dim_t)
            xs.append(np.asarray(bl, np.float32))
            ys.append(np.asarray(bh, np.float32))
        out.update({tag + "_low": low, tag + "_high": high, tag + "_aug_low": np.stack(xs), tag + "_aug_high": np.stack(ys),
                    tag + "_cfg": np.array([T, L, u, nframes, dmin, smin, smax, rot, flip, 2000 + seed, 3000 + seed], np.float64)})
        if dim_t > 1:
            # getinput of the 8x trainer on the same three-frame data: selectRandomTiles with the DEFAULT tile_t = 1 picks one
            # frame of the sequence (getRandomDatum :548-560: randrange(0, dim_t - tile_t), i.e. never the last one)
            for key, aug in (("single", False), ("single_aug", True)):
                rnd.seed(4000 + seed)
                np.random.seed(5000 + seed)
                xs, ys = [], []
                for call in range(3):
                    bl, bh = st.selectRandomTiles(6, isTraining=True, augment=aug)
                    xs.append(np.asarray(bl, np.float32))
                    ys.append(np.asarray(bh, np.float32))
                out.update({"%s_%s_low" % (tag, key): np.stack(xs), "%s_%s_high" % (tag, key): np.stack(ys)})
            out[tag + "_single_seeds"] = np.array([4000 + seed, 5000 + seed], np.float64)
    np.savez_compressed(os.path.join(HERE, "tileaugment.npz"), **out)
    print("tileaugment.npz:", {k: v.shape for k, v in out.items()})


def make_slice_fixtures():
    """The conv_slices path of FluidDataLoader.loadFiles (tools_wscale/fluiddataloader.py): the two `if self.conv_slices:`
    statements of loadFiles are lifted out of the method body by line number (:414-431 axis conversion / channel swap for
    x, :511-524 zoom + addAdjSlices + removeSlices + selectRandomSamples) and executed with the class's own helper methods
    on a stub `self`; the y branch of the axis conversion (:451-467) is the same code on fy."""
    import scipy.ndimage
    path = os.path.join(REF, "tools_wscale", "fluiddataloader.py")
    with open(path) as fh:
        tree = ast.parse(fh.read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FluidDataLoader"][0]
    load = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "loadFiles"][0]
    ifs = [n for n in ast.walk(load) if isinstance(n, ast.If) and isinstance(n.test, ast.Attribute) and n.test.attr == "conv_slices"]
    conv_x = [n for n in ifs if n.lineno in range(410, 420)][0]
    post = [n for n in ifs if n.lineno in range(508, 514)][0]

    def as_fn(name, args, stmt, ret):
        fn = ast.FunctionDef(name=name, args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=a) for a in args], kwonlyargs=[],
                                                           kw_defaults=[], defaults=[]),
                             body=[stmt, ast.parse("return " + ret).body[0]], decorator_list=[], type_params=[])
        return fn

    helpers = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("removeSlices", "addAdjSlices", "selectRandomSamples")]
    mod = ast.Module(body=helpers + [as_fn("conv_x", ["self", "fx"], conv_x, "fx"), as_fn("post", ["self", "fx", "fy"], post, "(fx, fy)")],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = dict(np=np, scipy=scipy, FDG_DTYPE=np.float32)
    exec(compile(mod, path, "exec"), ns)

    class Stub:
        pass

    for k in ("removeSlices", "addAdjSlices", "selectRandomSamples", "conv_x", "post"):
        setattr(Stub, k, ns[k])
    out = {}
    rng = np.random.default_rng(31)
    L, u = 6, 4
    cases = {"mode2": (0, [1, 1, 1, 1], [0.25, 1, 1, 1], 4, False), "mode3": (1, [1, 1, 1, 1], [1, 1, 1, 1], 4, False),
             "mode1_tempo": (2, [1, 1, 1, 1], [1, 1, 1, 1], 12, False), "adj": (0, [1, 1, 1, 1], [0.25, 1, 1, 1], 12, True)}
    for tag, (axis, sc, scy, C, adj) in cases.items():
        same = axis != 0  # refinement modes: x is already at the high resolution
        shp = (L * u,) * 3 if same else (L,) * 3
        fx = rng.random(shp + (C,), dtype=np.float32)
        fx[..., 0] *= (rng.random(shp) < 0.03) * 1.0
        fx[: shp[0] // 3, ..., 0] = 0.0
        fy = rng.random((L * u,) * 3 + (1,), dtype=np.float32)
        st = Stub()
        st.conv_slices, st.conv_axis, st.have_y_npz = True, axis, True
        st.axis_scaling, st.axis_scaling_y, st.add_adj_idcs = sc, scy, adj
        st.density_threshold, st.select_random = 0.002, 0.5
        x1 = st.conv_x(np.copy(fx))
        y1 = st.conv_x(np.copy(fy))
        np.random.seed(7)
        x2, y2 = st.post(np.copy(x1), np.copy(y1))
        out.update({tag + "_fx": fx, tag + "_fy": fy, tag + "_x": np.asarray(x2, np.float32), tag + "_y": np.asarray(y2, np.float32),
                    tag + "_cfg": np.array([axis, C, int(adj)] + sc + scy, np.float64)})
    np.savez_compressed(os.path.join(HERE, "slicedata.npz"), **out)
    print("slicedata.npz:", {k: v.shape for k, v in out.items()})


def make_schedule_fixtures():
    """The growing / blending / learning-rate-step schedule of the 8x trainer, traced by EXECUTING the reference's own
    statements: the initialisation block in front of the training loop (GAN/multipassGAN-8x.py:1884-1896), the initial
    `currentUpres` (:211-218) and, from the body of `for it in range(startingIter, trainingIterations)` (:1898), every
    statement that drives the schedule (:1905-1916 without the data re-loading, :1966-1982); copyAdamVariables / saveModel
    are recording stubs.  Output: one row (it, currentUpres, index, currBlendPer, lrgs, grew) per iteration."""
    path = os.path.join(REF, "GAN", "multipassGAN-8x.py")
    with open(path) as fh:
        tree = ast.parse(fh.read())
    loop = [n for n in ast.walk(tree) if isinstance(n, ast.For) and isinstance(n.target, ast.Name) and n.target.id == "it"
            and "trainingIterations" in ast.unparse(n.iter)][0]
    parent = [n for n in ast.walk(tree) if isinstance(n, ast.Try) and loop in n.body][0]
    sched_names = {"currBlendPer", "interpolate_Perc", "start_interpol", "interpol_c", "lrgs"}

    def assigns(n):
        return isinstance(n, ast.Assign) and all(isinstance(t, ast.Name) and t.id in sched_names for t in n.targets)

    init = [n for n in parent.body[:parent.body.index(loop)]
            if assigns(n) or (isinstance(n, ast.If) and "startingIter" in ast.unparse(n.test))]
    up0 = [n for n in tree.body if isinstance(n, ast.If) and ast.unparse(n.test) == "outputOnly"
           and "currentUpres" in ast.unparse(n)][0]
    keys = ("start_interpol", "interpolate_Perc", "currBlendPer", "decayLR")
    body = []
    for n in loop.body:
        if isinstance(n, ast.If) and any(k in ast.unparse(n.test) for k in keys) and not ast.unparse(n.test).startswith("0 and"):
            if "currentUpres < upRes" in ast.unparse(n.test):  # the growing event: drop the data re-loading and the print
                n.body = [m for m in n.body if not (isinstance(m, ast.If) and "upsampling_mode" in ast.unparse(m.test))
                          and not (isinstance(m, ast.Expr) and "print" in ast.unparse(m))]
            body.append(n)
        elif isinstance(n, ast.Assign) and ast.unparse(n.targets[0]) == "index":
            body.append(n)
    assert len(init) >= 6 and len(body) == 7, (len(init), len(body))
    body.append(ast.parse("trace.append((it, currentUpres, index, currBlendPer, lrgs, len(events)))").body[0])
    loop.body = body
    mod = ast.Module(body=[up0] + init + [loop], type_ignores=[])
    ast.fix_missing_locations(mod)
    code = compile(mod, path, "exec")
    out = {}
    cases = {"m2_s5": dict(upsampling_mode=2, stageIter=5, decayIter=7, startingIter=0, upRes=8, decayLR=True),
             "m2_s4_resume9": dict(upsampling_mode=2, stageIter=4, decayIter=3, startingIter=9, upRes=8, decayLR=True),
             "m2_s4_resume12": dict(upsampling_mode=2, stageIter=4, decayIter=3, startingIter=12, upRes=8, decayLR=False),
             "m2_s3_resume17": dict(upsampling_mode=2, stageIter=3, decayIter=4, startingIter=17, upRes=8, decayLR=True),
             "m1_s1": dict(upsampling_mode=1, stageIter=1, decayIter=9, startingIter=0, upRes=8, decayLR=True),
             "m1_s3": dict(upsampling_mode=1, stageIter=3, decayIter=5, startingIter=0, upRes=8, decayLR=True),
             "m2_u4_s3": dict(upsampling_mode=2, stageIter=3, decayIter=2, startingIter=0, upRes=4, decayLR=True)}
    import time as _time
    for tag, c in cases.items():
        events, trace = [], []
        ns = dict(math=math, time=_time, outputOnly=False, trace=trace, events=events,
                  trainingIterations=c["stageIter"] * 6 + c["decayIter"],                    # :166
                  copyAdamVariables=lambda u: events.append(("copy", u)), saveModel=lambda c_: events.append(("save", c_)), **c)
        exec(code, ns)
        out[tag] = np.array(trace, np.float64)
        out[tag + "_cfg"] = json.dumps(c)
        print(tag, len(trace), "iterations; growing events", events)
    # the 1-in-20 empty-density batches of getinput (:1527-1533): that one statement, executed on seeded numpy state
    gi = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "getinput"][0]
    stmt = [n for n in gi.body if isinstance(n, ast.If) and "randint" in ast.unparse(n.test)][0]
    zcode = compile(ast.fix_missing_locations(ast.Module(body=[stmt], type_ignores=[])), path, "exec")
    rng = np.random.default_rng(77)
    xs0 = rng.random((3, 1, 4, 4, 6), dtype=np.float32)
    ys0 = rng.random((3, 1, 8, 8, 1), dtype=np.float32)
    hits = []
    for seed in range(60):
        np.random.seed(seed)
        ns = dict(np=np, batch_xs=xs0.copy(), batch_ys=ys0.copy(), add_adj_idcs=1)
        exec(zcode, ns)
        if not np.array_equal(ns["batch_xs"], xs0):
            hits.append(seed)
            if len(hits) == 1:
                out["zero_seed"], out["zero_xs"], out["zero_ys"] = np.array(seed), ns["batch_xs"], ns["batch_ys"]
    out["zero_xs0"], out["zero_ys0"], out["zero_hits"] = xs0, ys0, np.array(hits)
    print("empty-density batches at seeds", hits)
    np.savez_compressed(os.path.join(HERE, "schedule8x.npz"), **out)


def make_tempo_fixtures():
    """getTempoinput's data side (tools_wscale/tilecreator_t.py): selectRandomTempoTiles (:1382-1413, the (sample, frame)
    re-ordering of three-frame tiles and the per-frame dt) and getSemiLagrPosBatch / gridInterpolBatch /
    getMACGridCenteredBatch (:1291-1378, the semi-Lagrangian re-sampling positions), executed as they are on a stub `self`
    whose selectRandomTiles hands back prepared three-frame tiles."""
    import scipy.ndimage
    path = os.path.join(REF, "tools_wscale", "tilecreator_t.py")
    code = ref_functions(path, ["gridInterpolBatch", "getMACGridCenteredBatch", "getSemiLagrPosBatch", "selectRandomTempoTiles"])
    ns = dict(np=np, scipy=scipy, DATA_KEY_LOW=0, DATA_KEY_HIGH=1, C_KEY_VELOCITY="v")
    exec(code, ns)
    rng = np.random.default_rng(41)
    T, u, C, B, n_t = 4, 4, 6, 2, 3
    S = T * u
    low = rng.standard_normal((B, 1, T, T, C * n_t)).astype(np.float32)
    high = rng.random((B, 1, S, S, n_t), dtype=np.float32)

    class Stub:
        tileSizeLow, tileSizeHigh, dim = [1, T, T], [1, S, S], 2
        c_lists = {0: {"v": [[1, 2, 3]]}}

        def selectRandomTiles(self, n, isTraining, augment, tile_t=1):
            assert n == B and tile_t == n_t
            return low.copy(), high.copy()

    xt, yt, pos = ns["selectRandomTempoTiles"](Stub(), B * n_t, True, False, n_t, 0.5)
    out = dict(low=low, high=high, xt=np.asarray(xt, np.float32), yt=np.asarray(yt, np.float32), pos=np.asarray(pos, np.float64),
               cfg=np.array([T, u, C, B, n_t, 0.5]))
    # positions alone, other sizes (incl. the no-interpolation branch cube_len_output == tile size)
    vel = rng.standard_normal((5, 1, 3, 3, 3)).astype(np.float32)
    dta = np.array([0.5, 0.0, -0.5, 0.25, -1.0], np.float32).reshape(-1, 1, 1, 1)
    out["vel"], out["dt"] = vel, dta.ravel()
    out["pos_up8"] = np.asarray(ns["getSemiLagrPosBatch"](vel, dta, 24), np.float64)
    out["pos_same"] = np.asarray(ns["getSemiLagrPosBatch"](vel, dta, 3), np.float64)
    # which frames FluidDataLoader loads for a data_fraction (tools_wscale/fluiddataloader.py:238-244, the "simple index range"
    # branch): the three statements that compute n, tf and filelist_index, executed for several ranges
    fpath = os.path.join(REF, "tools_wscale", "fluiddataloader.py")
    with open(fpath) as fh:
        ftree = ast.parse(fh.read())
    loops = [n for n in ast.walk(ftree) if isinstance(n, ast.For) and "self.filename_index_min + t * tf" in ast.unparse(n)]
    assert len(loops) == 1
    loop = loops[0]
    parent = [n for n in ast.walk(ftree) if isinstance(n, ast.If) and loop in n.orelse][0]
    pre = parent.orelse[:parent.orelse.index(loop)]
    loop.body = [loop.body[0], ast.parse("picked.append(filelist_index)").body[0]]
    fcode = compile(ast.fix_missing_locations(ast.Module(body=pre + [loop], type_ignores=[])), fpath, "exec")
    cases = [(0, 120, 0.08), (0, 120, 0.16), (3, 123, 0.08), (6, 126, 0.08), (0, 5, 1.0), (10, 50, 0.3), (0, 200, 0.01), (0, 7, 0.0)]
    rows = []
    for lo, hi, frac in cases:
        picked = []
        exec(fcode, dict(self=types.SimpleNamespace(filename_index_min=lo, filename_index_max=hi, data_fraction=frac), picked=picked))
        rows.append(picked)
    out["frac_cases"] = np.array(cases, np.float64)
    out["frac_picked"] = json.dumps(rows)
    print("frames per data_fraction:", rows[0], rows[4])
    np.savez_compressed(os.path.join(HERE, "tempotiles.npz"), **out)
    print("tempotiles.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    assert os.path.isdir(REF), "run in the authoring container (needs /root/reference)"
    if "slices" in (sys.argv[1:] or ["slices"]):
        make_slice_fixtures()
    which = sys.argv[1:] or ["pipeline", "nets", "tiles", "uni", "sampler", "augment"]
    if "augment" in which:
        make_augment_fixtures()
    if "sampler" in which:
        make_sampler_fixtures()
    if "pipeline" in which:
        make_pipeline_fixtures()
    if "nets" in which:
        make_net_fixtures()
    if "growdisc" in which:
        make_growdisc_fixtures()
    if "schedule" in which:
        make_schedule_fixtures()
    if "tempo" in which:
        make_tempo_fixtures()
    if "tiles" in which:
        make_tile_fixtures()
    if "uni" in which:
        make_uni_fixtures()
