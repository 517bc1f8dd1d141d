"""Oracle + host graph builders vs the golden vectors in tests/golden/ (CPU).

The vectors were produced by tests/golden/make_golden.py EXECUTING THE REFERENCE'S OWN PYTHON
(generate3DUniForNewNetwork of multipassGAN-out.py / multipassGAN-4x.py with stand-in row functions;
tools_wscale/GAN.py + growing_gen / gen_resnet / disc_binclass on a numpy TF1 shim).  They pin
  * the volume pipeline: bit-exact,
  * layer wiring + variable names/shapes of every generator (a1-a13) and the discriminator (a20): the
    fp64 oracle must reproduce the reference-code output to 1e-9 (independent conv/resize restatements).
"""
import json
import os
import sys
import zlib

import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import graph as G, networks as N, pipeline as P, weights as W
from oracle import gan as og, networks as on, pipeline as op
from oracle_nets import oracle_gen_resnet, oracle_growing_gen

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import standin  # noqa: E402


@pytest.fixture(scope="module")
def pipe():
    return np.load(os.path.join(GOLD, "pipeline.npz"))


@pytest.fixture(scope="module")
def nets():
    return np.load(os.path.join(GOLD, "nets.npz"))


# ------------------------------------------------------------------------------ volume pipeline
OUT_CASES = [(0, (1, 2), True), (0, (1, 2, 3), True), (1, (1, 2, 3), False), (3, (1, 2), True), (2, (1,), True),
             (0, (1,), False)]


@pytest.mark.parametrize("ta,which,adj", OUT_CASES)
def test_out_pipeline_matches_reference_code(pipe, ta, which, adj):
    x, L, u = pipe["x"], int(pipe["L"]), int(pipe["u"])
    key = "out_ta%d_n%s_adj%d" % (ta, "".join(map(str, which)), int(adj))
    log = {}

    def tap(name, fn):
        def run(*rows):
            log.setdefault(name, []).append(np.array(rows[0]))
            return fn(*rows)
        return run

    n1 = tap("sampler_x", lambda r: standin.net_first(r, L, u, 6 if adj else 4)) if 1 in which else None
    n2 = tap("sampler_2_x", lambda r, y: standin.net_refine(r, y, L, u, 2)) if 2 in which else None
    n3 = tap("sampler_3_x", lambda r, y: standin.net_refine(r, y, L, u, 3)) if 3 in which else None
    vol = op.out_generate3d(x, u, n1, n2, n3, transposeAxis=ta, add_adj_idcs1=adj, threshold=True)
    np.testing.assert_array_equal(vol, pipe[key + "_vol"])
    for name, chunks in log.items():
        got = np.concatenate([c.reshape(-1) for c in chunks])
        np.testing.assert_array_equal(got, pipe[key + "_" + name].reshape(-1))


def test_4x_passes_match_reference_code(pipe):
    x, L, u = pipe["x"], int(pipe["L"]), int(pipe["u"])
    S = L * u
    p1 = op.apply_4x_pass(lambda r: standin.net_first(r, L, u, 4), u, 2, x)
    np.testing.assert_array_equal(p1, pipe["x4_p1_vol"])
    vel = x[..., 1:4] * u
    for mode, key in ((1, "x4_p2"), (3, "x4_p3")):
        feeds = []

        def net(r):
            feeds.append(np.array(r))
            return standin.net_fullres(r, S)

        got = op.apply_4x_pass(net, u, mode, vel, x_2=p1[..., None])
        np.testing.assert_array_equal(got, pipe[key + "_vol"])
        np.testing.assert_array_equal(np.concatenate(feeds).reshape(-1), pipe[key + "_feed"].reshape(-1))


# ------------------------------------------------------------------------------ networks
def _wsum(values, names):
    return float(sum(np.float64(zlib.crc32(np.ascontiguousarray(values[n]).tobytes())) for n in names))


def _check_vars(graph, blob, weights, wsum):
    want = {k: tuple(v) for k, v in json.loads(str(blob))}
    have = {v.name: v.shape for v in graph.variables.values()}
    assert have == want  # names (checkpoint keys) AND shapes, exactly what the reference code requested
    assert _wsum(weights, list(want)) == float(wsum)  # same deterministic values as the fixture run


def _close(got, ref, tol=1e-9):
    ref = np.asarray(ref, np.float64)
    err = np.abs(np.asarray(got, np.float64) - ref).max()
    assert err <= tol * max(1.0, np.abs(ref).max()), err


@pytest.mark.parametrize("tag", ["out_net1", "out_net1_u8", "out_net2", "out_net3"])
def test_growing_gen_matches_reference_code(nets, tag):
    cfg = json.loads(str(nets[tag + "_cfg"]))
    s = cfg["spec"]
    spec = P.NetSpec(use_res_net=s["use_res_net"], add_adj_idcs=s["add_adj_idcs"], startFms=s["startFms"],
                     maxFms=s["maxFms"], filterSize=s["filterSize"], first_nn_arch=s["first_nn_arch"])
    L, u, idx = cfg["L"], cfg["u"], s["idx"]
    G.reset_default_graph()
    P.build_out_graph(idx, spec, N.config_out(L, upRes=u))
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), cfg["seed"]), cfg["seed"])
    _check_vars(G.get_default_graph(), nets[tag + "_vars"], w, nets[tag + "_wsum"])
    y = nets[tag + "_y"] if tag + "_y" in nets else None
    got = oracle_growing_gen(w, idx, spec, L, upRes=u)(nets[tag + "_x"], y)
    _close(got, nets[tag + "_out"])


@pytest.mark.parametrize("tag", ["x4_mode2", "x4_mode1", "x4_mode2_nobn"])
def test_gen_resnet_and_disc_match_reference_code(nets, tag):
    cfg = json.loads(str(nets[tag + "_cfg"]))
    L, u, mode, bn, seed = cfg["L"], cfg["u"], cfg["mode"], cfg["batch_norm"], cfg["seed"]
    G.reset_default_graph()
    c = N.config_4x(L, upRes=u, upsampling_mode=mode, batch_norm=bn)
    n_in = (L * L if mode == 2 else (L * u) ** 2) * 4
    N.gen_resnet(G.placeholder([None, n_in], "x"), c)
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), seed), seed)
    _check_vars(G.get_default_graph(), nets[tag + "_vars"], w, nets[tag + "_wsum"])
    got = oracle_gen_resnet(w, L, mode, upRes=u, batch_norm=bn)(nets[tag + "_x"])
    _close(got, nets[tag + "_out"])
    if tag + "_disc_logits" not in nets:
        return
    # discriminator (inference-mode BN): the oracle creates its variables on demand with the same
    # name-seeded initialiser, then they are compared with what the reference code requested
    store = og.VarStore(seed=seed, values=w)
    ctx = og.Context(store, torch.float64)
    ocfg = on.make_cfg_4x(L, upRes=u, upsampling_mode=mode, batch_norm=bn)
    # BN variables of the disc need the fixture's randomised statistics: create defaults, randomise, re-run
    on.disc_binclass(torch.as_tensor(nets[tag + "_x"]).double(), torch.as_tensor(nets[tag + "_disc_y"]).double(), ctx,
                     ocfg, train=False, use_batch_norm=bn)
    store.values = W.randomize_bn_stats(store.values, seed)
    want = {k: tuple(v) for k, v in json.loads(str(nets[tag + "_disc_vars"]))}
    have = {n: tuple(store.values[n].shape) for n in store.values if n.startswith("discriminator/")}
    assert have == want
    assert _wsum(store.values, list(w) + list(want)) == float(nets[tag + "_disc_wsum"])
    ctx = og.Context(store, torch.float64)
    d = on.disc_binclass(torch.as_tensor(nets[tag + "_x"]).double(), torch.as_tensor(nets[tag + "_disc_y"]).double(),
                         ctx, ocfg, train=False, use_batch_norm=bn)
    _close(d[0].numpy(), nets[tag + "_disc_logits"])
    for i in range(1, 5):
        _close(d[i].numpy(), nets[tag + "_disc_d%d" % i])


# ------------------------------------------------------------------------------ tiles (a19)
def test_tile_cut_and_stitch_match_reference_code():
    from oracle import tiles as otl
    g = np.load(os.path.join(GOLD, "tiles.npz"))
    frame = g["frame"]
    for tag, (tile, stride, pad) in {"reg": ([1, 8, 8], -1, 0), "ovl": ([1, 12, 16], 4, 0), "ovl2": ([1, 8, 10], 6, 0)}.items():
        np.testing.assert_array_equal(otl.create_tiles(frame, tile, stride, pad), g["tiles_" + tag])
    # edge padding (not pinned, see make_golden.py): the tile's own border pixels are replicated in y and x
    t = otl.create_tiles(frame, [1, 8, 10], 6, 2)
    assert t.shape[1:] == (1, 12, 14, 3) and np.array_equal(t[:, :, 2:-2, 2:-2], g["tiles_ovl2"])
    assert np.array_equal(t[:, :, 0, 2:-2], g["tiles_ovl2"][:, :, 0]) and np.array_equal(t[:, :, 2:-2, -1], g["tiles_ovl2"][:, :, :, -1])
    np.testing.assert_array_equal(otl.concat_tiles(g["stitch_in"], [1, 2, 3], (0, 2, 2, 0)), g["stitch_b2"])
    np.testing.assert_array_equal(otl.concat_tiles(g["stitch_in"], [1, 2, 3]), g["stitch_b0"])
