"""CUDA path vs the golden vectors produced by executing the reference's own Python (tests/golden/).

  * generators through the C ABI (fp32 path <= 1e-4, fp16 tensor-core path rel-L2 <= 5e-3 / max-abs <= 2e-2)
  * the device-resident volume pipelines (slice assembler, slice batching, inter-pass transposes,
    threshold) with the golden run's stand-in row functions substituted for the networks.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import engine, graph as G, networks as N, pipeline as P, weights as W
from oracle_nets import err_stats

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import standin  # noqa: E402

TOL = {"fp32": (1e-4, 1e-4), "fp16": (5e-3, 2e-2)}


@pytest.fixture(scope="module")
def pipe():
    return np.load(os.path.join(GOLD, "pipeline.npz"))


@pytest.fixture(scope="module")
def nets():
    return np.load(os.path.join(GOLD, "nets.npz"))


def _check(name, got, ref, precision):
    st = err_stats(got, ref)
    rel_tol, abs_tol = TOL[precision]
    print("%s [%s] rel_l2=%.3e max_abs=%.3e ref_max=%.3f" % (name, precision, st["rel_l2"], st["max_abs"], st["ref_max"]))
    assert np.isfinite(got).all()
    assert st["rel_l2"] <= rel_tol and st["max_abs"] <= abs_tol * max(1.0, st["ref_max"]), (name, precision, st)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("tag", ["out_net1", "out_net1_u8", "out_net2", "out_net3"])
def test_growing_gen_vs_reference_code(nets, tag, precision):
    cfg = json.loads(str(nets[tag + "_cfg"]))
    s = cfg["spec"]
    spec = P.NetSpec(use_res_net=s["use_res_net"], add_adj_idcs=s["add_adj_idcs"], startFms=s["startFms"],
                     maxFms=s["maxFms"], filterSize=s["filterSize"], first_nn_arch=s["first_nn_arch"])
    G.reset_default_graph()
    out = P.build_out_graph(s["idx"], spec, N.config_out(cfg["L"], upRes=cfg["u"]))
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), cfg["seed"]), cfg["seed"])
    x = nets[tag + "_x"]
    feeds = {"x": torch.from_numpy(x).cuda()}
    if tag + "_y" in nets:
        feeds["y"] = torch.from_numpy(nets[tag + "_y"]).cuda()
    net = engine.CompiledNet(out, w, x.shape[0], precision=precision)
    y = net.run(feeds).float().cpu().numpy()
    _check(tag, y, nets[tag + "_out"], precision)
    net.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("tag", ["x4_mode2", "x4_mode1", "x4_mode2_nobn"])
def test_gen_resnet_vs_reference_code(nets, tag, precision):
    cfg = json.loads(str(nets[tag + "_cfg"]))
    L, u, mode, bn, seed = cfg["L"], cfg["u"], cfg["mode"], cfg["batch_norm"], cfg["seed"]
    G.reset_default_graph()
    n_in = (L * L if mode == 2 else (L * u) ** 2) * 4
    out = N.gen_resnet(G.placeholder([None, n_in], "x"), N.config_4x(L, upRes=u, upsampling_mode=mode, batch_norm=bn))
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), seed), seed)
    x = nets[tag + "_x"]
    net = engine.CompiledNet(out, w, x.shape[0], precision=precision)
    y = net.run({"x": torch.from_numpy(x).cuda()}).float().cpu().numpy()
    _check(tag, y, nets[tag + "_out"], precision)
    net.close()


# ------------------------------------------------------------------------------ pipelines with stand-in networks
def _span(t, rows, cols):
    """[rows, cols] view starting at the first element of `t` (the pipelines hand the networks the pointer
    of the first slice of a batch, like the C ABI does)."""
    return t.new_empty(0).set_(t.untyped_storage(), t.storage_offset(), (rows, cols), (cols, 1))


class _StubNet:
    """Test double for engine.CompiledNet: runs a numpy row function on host copies of the feeds."""

    def __init__(self, fn, batch, n_out):
        self.fn, self.batch, self.n_out = fn, batch, n_out
        self.net = self
        self.flops, self.launches = 0.0, 0

    def run(self, feeds, out=None, stream=None):
        torch.cuda.synchronize()
        x = feeds["x"].detach().cpu().numpy().reshape(self.batch, -1)
        y = _span(feeds["y"], self.batch, self.n_out).cpu().numpy() if "y" in feeds else None
        res = self.fn(x, y)
        _span(out, self.batch, self.n_out).copy_(torch.from_numpy(np.ascontiguousarray(res, dtype=np.float32)))
        torch.cuda.synchronize()
        return out


OUT_CASES = [(0, (1, 2), True), (0, (1, 2, 3), True), (1, (1, 2, 3), False), (0, (1,), False)]


@pytest.mark.parametrize("ta,which,adj", OUT_CASES)
def test_out_pipeline_vs_reference_code(pipe, ta, which, adj):
    x, L, u = pipe["x"], int(pipe["L"]), int(pipe["u"])
    S = L * u
    specs = {1: P.NetSpec(True, adj, 16, 16, 3, True), 2: P.NetSpec(True, False, 16, 16, 3),
             3: P.NetSpec(False, False, 16, 8, 3)}
    w = P.make_weights_out(L, 1, upRes=u, specs=specs, nets=which)
    mp = P.MultiPassOut(L, w, upRes=u, specs=specs, precision="fp32", transposeAxis=ta, batches=(8, 2, 2))
    fns = {1: lambda r, y: standin.net_first(r, L, u, 6 if adj else 4),
           2: lambda r, y: standin.net_refine(r, y, L, u, 2), 3: lambda r, y: standin.net_refine(r, y, L, u, 3)}
    for idx in which:
        mp.passes[idx]["net"] = _StubNet(fns[idx], mp.passes[idx]["batch"], S * S)
    got = mp(x).cpu().numpy()
    key = "out_ta%d_n%s_adj%d" % (ta, "".join(map(str, which)), int(adj))
    ref = pipe[key + "_vol"]
    # the slice assembler lerps in fp32 where scipy.ndimage.zoom works in double: <= a few ulp per element
    _check(key, got, ref, "fp32")
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_4x_two_pass_pipeline_vs_reference_code(pipe):
    x, L, u = pipe["x"], int(pipe["L"]), int(pipe["u"])
    S = L * u
    w1, w2 = P.make_weights_4x(L, 1, upRes=u)
    mp = P.MultiPass4x(L, w1, w2, upRes=u, precision="fp32", batch=8)
    mp.p1 = _StubNet(lambda r, y: standin.net_first(r, L, u, 4), mp.batch, S * S)
    mp.p2 = _StubNet(lambda r, y: standin.net_fullres(r, S), mp.batch, S * S)
    got = mp(x).cpu().numpy()
    ref = pipe["x4_p2_vol"]
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    p1 = mp.pass1_only(x).cpu().numpy()
    assert np.abs(p1 - pipe["x4_p1_vol"]).max() <= 2e-5 * max(1.0, np.abs(pipe["x4_p1_vol"]).max())


# ------------------------------------------------------------------------------ tiles (a19)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_tile_cut_and_stitch_vs_reference_code(dtype):
    from mpgan_b200 import capi
    g = np.load(os.path.join(GOLD, "tiles.npz"))
    h = capi.default_handle(0)
    frame = torch.from_numpy(g["frame"]).to(dtype).cuda()  # [1, 24, 40, 3]
    eb = frame.element_size()
    from oracle import tiles as otl
    for tag, (th, tw, stride, pad) in {"reg": (8, 8, -1, 0), "ovl": (12, 16, 4, 0), "ovl2": (8, 10, 6, 0), "pad": (8, 10, 6, 2)}.items():
        if tag == "pad":  # edge padding: checked against the oracle (the reference branch raises, see make_golden.py)
            ref = torch.from_numpy(otl.create_tiles(g["frame"], [1, th, tw], stride, pad)).to(dtype)
        else:
            ref = torch.from_numpy(g["tiles_" + tag]).to(dtype)  # [tiles, 1, th, tw, 3]
        ty, tx = capi.tiles_count(24, th, stride), capi.tiles_count(40, tw, stride)
        assert ty * tx == ref.shape[0]
        out = torch.empty((ty * tx, th + 2 * pad, tw + 2 * pad, 3), dtype=dtype, device="cuda")
        capi.tiles_cut(h, frame, out, 1, 24, 40, 3, eb, th, tw, stride, stride, pad)
        assert torch.equal(out.cpu(), ref[:, 0])
    tiles = torch.from_numpy(g["stitch_in"]).to(dtype).cuda()  # [6, 1, 12, 16, 3]
    for border, key in ((2, "stitch_b2"), (0, "stitch_b0")):
        ref = torch.from_numpy(g[key]).to(dtype)
        out = torch.empty(tuple(ref.shape), dtype=dtype, device="cuda")
        capi.tiles_stitch(h, tiles, out, 1, 2, 3, 12, 16, 3, eb, border)
        assert torch.equal(out.cpu(), ref)
    with pytest.raises(capi.MpgError):
        capi.tiles_cut(h, frame, frame, 1, 24, 40, 3, eb, 64, 8)  # tile larger than the frame (TilecreatorError role)
