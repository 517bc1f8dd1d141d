"""GPU parity: whole generators through the drop-in layer API + engine vs the fp64 oracle.

Tolerances (BASELINE.json north_star): 16-bit path max-abs <= 2e-2 (scaled by max|ref| because
random-init outputs are not confined to [0,1]) and rel-L2 <= 5e-3; fp32 path <= 1e-4.
"""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import engine, graph as G, networks as N, pipeline as P, weights as W
from oracle_nets import err_stats, oracle_gen_resnet, oracle_growing_gen

pytestmark = pytest.mark.gpu

TOL = {"fp32": (1e-4, 1e-4), "fp16": (5e-3, 2e-2), "bf16": (5e-2, 3e-1)}  # (rel_l2, max_abs/scale); bf16: opt-in, see test_pipeline_gpu


def _check(name, got, ref, precision):
    st = err_stats(got, ref)
    rel_tol, abs_tol = TOL[precision]
    scale = max(1.0, st["ref_max"])
    print("%s [%s] rel_l2=%.3e max_abs=%.3e ref_max=%.3f" % (name, precision, st["rel_l2"], st["max_abs"], st["ref_max"]))
    assert np.isfinite(got).all()
    assert st["rel_l2"] <= rel_tol, (name, precision, st)
    assert st["max_abs"] <= abs_tol * scale, (name, precision, st)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("mode,L", [(2, 16), (1, 8)])
def test_gen_resnet(precision, mode, L):
    B, u = 4, 4
    G.reset_default_graph()
    cfg = N.config_4x(L, upRes=u, upsampling_mode=mode)
    n_in = L * L * 4 if mode == 2 else (L * u) ** 2 * 4
    out = N.gen_resnet(G.placeholder([None, n_in], "x"), cfg)
    w = W.randomize_bn_stats(W.init_graph_variables(G.get_default_graph(), 3), 3)
    rng = np.random.default_rng(0)
    x = (rng.random((B, n_in), dtype=np.float32) * np.tile([1, .5, .5, .5], n_in // 4)).astype(np.float32)
    net = engine.CompiledNet(out, w, B, precision=precision)
    y = net.run({"x": torch.from_numpy(x).cuda()}).float().cpu().numpy()
    ref = oracle_gen_resnet(w, L, mode)(x)
    _check("gen_resnet mode %d" % mode, y, ref, precision)
    net.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("idx", [1, 2])
def test_growing_gen_8x(precision, idx):
    L, u = 8, 8
    S = L * u
    spec = P.SHIPPED_8X[idx]
    B = 2
    w = P.make_weights_out(L, 5, upRes=u, nets=(idx,))[idx]
    G.reset_default_graph()
    cfg = N.config_out(L, upRes=u)
    out = P.build_out_graph(idx, spec, cfg)
    rng = np.random.default_rng(1)
    cin = 6 if idx == 1 else 4
    x = rng.random((B, L * L * cin), dtype=np.float32)
    feeds = {"x": torch.from_numpy(x).cuda()}
    yrows = None
    if idx == 2:
        yrows = rng.random((B, S * S), dtype=np.float32)
        feeds["y"] = torch.from_numpy(yrows).cuda()
    net = engine.CompiledNet(out, w, B, precision=precision)
    y = net.run(feeds).float().cpu().numpy()
    ref = oracle_growing_gen(w, idx, spec, L, upRes=u)(x, yrows)
    _check("growing_gen net%d" % idx, y, ref, precision)
    net.close()


def test_growing_gen_plain_chain_net3():
    """use_res_net 0 branch (third network as configured, lrelu + pixel_norm conv chain)."""
    L, u, idx, B = 8, 4, 3, 2
    S = L * u
    specs = {3: P.NetSpec(use_res_net=False, startFms=64, maxFms=32, filterSize=5)}
    w = P.make_weights_out(L, 6, upRes=u, specs=specs, nets=(3,))[3]
    G.reset_default_graph()
    out = P.build_out_graph(3, specs[3], N.config_out(L, upRes=u))
    rng = np.random.default_rng(2)
    x = rng.random((B, L * L * 4), dtype=np.float32)
    yr = rng.random((B, S * S), dtype=np.float32)
    net = engine.CompiledNet(out, w, B, precision="fp32")
    y = net.run({"x": torch.from_numpy(x).cuda(), "y": torch.from_numpy(yr).cuda()}).cpu().numpy()
    ref = oracle_growing_gen(w, idx, specs[3], L, upRes=u)(x, yr)
    _check("growing_gen net3 chain", y, ref, "fp32")


def test_fp16_range_check_raises_instead_of_clipping():
    """Every 16-bit store saturates silently (cvt.rn.satfinite). In validation mode (range_check=True) a generator whose
    intermediate activations leave the fp16 range must RAISE, with the offending layers named; the same weights pass in
    fp32, and sane weights pass the check in fp16."""
    from mpgan_b200 import capi
    L, B, u = 8, 2, 4
    G.reset_default_graph()
    cfg = N.config_4x(L, upRes=u, upsampling_mode=2)
    out = N.gen_resnet(G.placeholder([None, L * L * 4], "x"), cfg)
    w = W.init_graph_variables(G.get_default_graph(), 3)
    x = torch.from_numpy(np.random.default_rng(0).random((B, L * L * 4), dtype=np.float32)).cuda()
    net = engine.CompiledNet(out, w, B, precision="fp16", range_check=True)
    net.run({"x": x})  # random-init activations peak near 13: no saturation
    assert net.saturated() == {}
    net.close()
    big = dict(w)
    big["generator/g_cA1/weight"] = w["generator/g_cA1/weight"] * 3e4  # ru2 conv A now produces values beyond 65504
    net = engine.CompiledNet(out, big, B, precision="fp16", range_check=True)
    with pytest.raises(capi.MpgRangeError) as ei:
        net.run({"x": x})
    assert any("g_cA1" in k for k in ei.value.counts), ei.value.counts
    net.close()
    # without the check the same run returns (clipped) finite data: this is the silent behaviour the mode exists for
    net = engine.CompiledNet(out, big, B, precision="fp16")
    y = net.run({"x": x})
    assert torch.isfinite(y).all()
    net.close()
    # the fused head keeps its intermediate in shared memory: its in-kernel counter must see a saturating conv A
    big0 = dict(w)
    big0["generator/g_cA0/weight"] = w["generator/g_cA0/weight"] * 1e5
    net = engine.CompiledNet(out, big0, B, precision="fp16", range_check=True)
    with pytest.raises(capi.MpgRangeError) as ei:
        net.run({"x": x})
    assert any("g_cA0" in k for k in ei.value.counts), ei.value.counts
    net.close()
    net32 = engine.CompiledNet(out, big, B, precision="fp32", range_check=True)
    net32.run({"x": x})
    net32.close()
