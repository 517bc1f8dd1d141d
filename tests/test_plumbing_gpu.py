"""GPU parity of the bandwidth kernels against numpy / scipy / the oracle op restatements (bit-exact
for pure data movement, 1e-6 for fp32 interpolation)."""
import numpy as np
import pytest
import torch

import mpgan_b200  # noqa: F401
from mpgan_b200 import capi
from oracle import pipeline as op
from oracle import tf_ops

pytestmark = pytest.mark.gpu


def H():
    return capi.default_handle(0)


@pytest.mark.parametrize("perm", [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)])
@pytest.mark.parametrize("dims", [(8, 8, 8), (5, 33, 70), (64, 64, 64), (1, 37, 2), (12, 68, 100), (4, 132, 60)])
def test_transpose3d_exact(perm, dims):
    rng = np.random.default_rng(0)
    a = rng.standard_normal(dims).astype(np.float32)
    src = torch.from_numpy(a).cuda()
    dst = torch.full(tuple(dims[p] for p in perm), float("nan"), device="cuda")
    capi.transpose3d(H(), src, dst, dims, perm)
    torch.cuda.synchronize()
    assert np.array_equal(dst.cpu().numpy(), a.transpose(perm))


def test_transpose3d_threshold_and_threshold():
    rng = np.random.default_rng(1)
    a = (rng.random((16, 24, 40)).astype(np.float32) * 0.002)
    src = torch.from_numpy(a).cuda()
    dst = torch.empty((40, 16, 24), device="cuda")
    capi.transpose3d(H(), src, dst, a.shape, (2, 0, 1), 0.0005)
    ref = a.transpose(2, 0, 1).copy()
    ref[ref < 0.0005] = 0
    assert np.array_equal(dst.cpu().numpy(), ref)
    v = torch.from_numpy(a).cuda()
    capi.threshold(H(), v, a.size, 0.0005)
    r2 = a.copy()
    r2[r2 < 0.0005] = 0
    assert np.array_equal(v.cpu().numpy(), r2)


@pytest.mark.parametrize("ta", [0, 1, 2, 3])
@pytest.mark.parametrize("adj", [False, True])
def test_slice_assemble_pass1_matches_reference_numpy(ta, adj):
    """GAN/multipassGAN-out.py:398-436 restated in oracle.pipeline.out_pass1_input."""
    from mpgan_b200.pipeline import _PASS_GEOM
    L, u, C = 6, 4, 4
    S = L * u
    rng = np.random.default_rng(2)
    x = rng.standard_normal((L, L, L, C)).astype(np.float32)
    ref = op.out_pass1_input(x, L, S, u, C, ta, adj)
    axis_of, chans = _PASS_GEOM[1][ta]
    cout = C + (2 if adj else 0)
    d = capi.make_assemble_desc((L, L, L), C, axis_of, (u, 1, 1), chans, None, add_adj=adj, out_dtype=capi.F32,
                                out_cstride=cout)
    out = torch.full((S, L, L, cout), float("nan"), device="cuda")
    capi.slice_assemble(H(), d, torch.from_numpy(x).cuda(), None, 0, S, out)
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("ta", [0, 1, 2, 3])
def test_slice_assemble_pass2_matches_reference_numpy(ta):
    from mpgan_b200.pipeline import _PASS_GEOM
    L, u, C = 5, 4, 4
    S = L * u
    x = np.random.default_rng(3).standard_normal((L, L, L, C)).astype(np.float32)
    ref = op.out_pass2_input(x, L, S, u, C, ta)
    axis_of, chans = _PASS_GEOM[2][ta]
    d = capi.make_assemble_desc((L, L, L), C, axis_of, (u, 1, 1), chans, None, out_dtype=capi.F32, out_cstride=C)
    out = torch.empty((S, L, L, C), device="cuda")
    # assembled in two ragged chunks to exercise slice0/count
    capi.slice_assemble(H(), d, torch.from_numpy(x).cuda(), None, 0, 7, out)
    capi.slice_assemble(H(), d, torch.from_numpy(x).cuda(), None, 7, S - 7, out[7])
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_slice_assemble_4x_pass2_trilinear_with_density():
    """GAN/multipassGAN-4x.py:1095,1113-1119: trilinear zoom of vel*u + first-pass density, (d,vy,vz,vx)."""
    L, u = 6, 4  # S = 24: a multiple of the reference's batch of 8 (App. D.4)
    S = L * u
    rng = np.random.default_rng(4)
    x = rng.standard_normal((L, L, L, 4)).astype(np.float32)
    dens = rng.random((S, S, S, 1)).astype(np.float32)
    captured = {}

    def net(rows):
        captured.setdefault("rows", []).append(np.array(rows))
        return np.zeros((rows.shape[0], S * S), np.float32)

    op.apply_4x_pass(net, u, 1, x[..., 1:4] * u, x_2=dens)
    ref = np.concatenate(captured["rows"]).reshape(-1, S, S, 4)
    n_ref = ref.shape[0]
    d = capi.make_assemble_desc((L, L, L), 4, (2, 0, 1), (u, u, u), (2, 3, 1), (u, u, u), out_dtype=capi.F32,
                                out_cstride=4)
    dens_t = np.ascontiguousarray(dens[..., 0].transpose(2, 0, 1))  # [X, Z, Y]
    out = torch.empty((S, S, S, 4), device="cuda")
    capi.slice_assemble(H(), d, torch.from_numpy(x).cuda(), torch.from_numpy(dens_t).cuda(), 0, S, out)
    got = out.cpu().numpy()[:n_ref]
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("out_dtype", ["f32", "bf16", "f16"])
def test_pack_channels(out_dtype):
    n, h, w = 2, 12, 20
    rng = np.random.default_rng(5)
    a = rng.standard_normal((n, h // 4, w // 4, 4)).astype(np.float32)  # nearest x4
    b = rng.standard_normal((n, h, w, 8)).astype(np.float32)
    tb = torch.from_numpy(b).cuda().to(torch.bfloat16)
    code = {"f32": capi.F32, "bf16": capi.BF16, "f16": capi.F16}[out_dtype]
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[out_dtype]
    out = torch.full((n, h, w, 8), float("nan"), dtype=tdt, device="cuda")
    capi.pack_channels(H(), [(tb, capi.BF16, 8, 2, 1, 1, 1), (torch.from_numpy(a).cuda(), capi.F32, 4, 0, 4, 4, 4)],
                       out, code, 8, n, h, w)
    ref = np.zeros((n, h, w, 8), np.float32)
    ref[..., 0] = tb.float().cpu().numpy()[..., 2]
    ref[..., 1:5] = a.repeat(4, axis=1).repeat(4, axis=2)
    ref_t = torch.from_numpy(ref).to(tdt).float().numpy()
    assert np.array_equal(out.float().cpu().numpy(), ref_t)


@pytest.mark.parametrize("factor", [2, 4, 8])
def test_dens_residual_bicubic_matches_oracle_tf1_resize(factor):
    n, L = 2, 9
    S = L * factor
    rng = np.random.default_rng(6)
    x = rng.random((n, L, L, 6)).astype(np.float32)
    dens = rng.standard_normal((n, S, S)).astype(np.float32)
    plan = capi.BicubicPlan(H(), L, L, S, S)
    out = torch.empty((n, S, S), device="cuda")
    capi.dens_residual(H(), torch.from_numpy(dens).cuda(), torch.from_numpy(x).cuda(), capi.F32, 6, 0, 2, plan, n, S, S,
                       L, L, out)
    ref = dens + tf_ops.resize_bicubic_tf1(torch.from_numpy(x[..., 0:1]), S, S).numpy()[..., 0]
    assert np.abs(out.cpu().numpy() - ref).max() <= 2e-6
    out0 = torch.empty((n, L, L), device="cuda")
    d0 = rng.standard_normal((n, L, L)).astype(np.float32)
    capi.dens_residual(H(), torch.from_numpy(d0).cuda(), torch.from_numpy(x).cuda(), capi.F32, 6, 3, 0, None, n, L, L, L,
                       L, out0)
    assert np.array_equal(out0.cpu().numpy(), d0 + x[..., 3])


def test_error_paths_raise():
    with pytest.raises(capi.MpgError):
        capi.transpose3d(H(), torch.empty(8, device="cuda"), torch.empty(8, device="cuda"), (2, 2, 2), (0, 0, 1))
    with pytest.raises(capi.MpgError):
        capi.ConvPlan(H(), 1, 8, 8, [np.zeros((3, 3, 4, 8), np.float32)], [4], 8, 8, force_kind=1)  # cstride % 8


def test_tiles_overlap_cut_and_stitch_roundtrip():
    """mpg_tiles_cut(stride = tile - 2b) followed by mpg_tiles_stitch_overlap(border b) is the identity on a frame of
    t*(tile - 2b) + 2b pixels; checked against numpy slicing for the cut and exact equality for the round trip."""
    from mpgan_b200 import capi
    h = capi.default_handle(0)
    rng = np.random.default_rng(4)
    n, c, tile, b, ty, tx = 3, 2, 12, 3, 4, 2
    core = tile - 2 * b
    H, W = ty * core + 2 * b, tx * core + 2 * b
    x = rng.standard_normal((n, H, W, c)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    tiles = torch.empty((n * ty * tx, tile, tile, c), device="cuda")
    capi.tiles_cut(h, xd, tiles, n, H, W, c, 4, tile, tile, core, core, 0, 0)
    want = np.stack([x[i, iy * core:iy * core + tile, ix * core:ix * core + tile] for i in range(n) for iy in range(ty)
                     for ix in range(tx)])
    assert np.array_equal(tiles.cpu().numpy(), want)
    back = torch.full((n, H, W, c), float("nan"), device="cuda")
    capi.tiles_stitch_overlap(h, tiles, back, n, ty, tx, tile, tile, c, 4, b, 0)
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), x)
    # a tile-local change inside the cropped band of an interior edge must not reach the output
    t2 = tiles.clone()
    t2[0, :, tile - 1, :] = 1e9  # right edge column of tile (0,0): belongs to tile (0,1)'s kept region
    capi.tiles_stitch_overlap(h, t2, back, n, ty, tx, tile, tile, c, 4, b, 0)
    torch.cuda.synchronize()
    assert np.array_equal(back.cpu().numpy(), x)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("split,perm", [(2, (2, 0, 1)), (2, (2, 1, 0)), (1, (1, 2, 0))])
def test_reslab_p2p_kernel_matches_numpy(world, split, perm):
    """mpg_reslab_p2p (fused transpose + stores into the owning rank's slab) with all 'ranks' emulated on one GPU: the
    peer pointers are G local buffers. Output slab h must be permute(full[:, :, h-range] or full[:, h-range, :], perm),
    exactly, with the threshold applied (the three (split, perm) pairs are the ones the pipelines use)."""
    from mpgan_b200 import capi
    h = capi.default_handle(0)
    S = 96
    per = S // world
    rng = np.random.default_rng(11)
    full = (rng.random((S, S, S), dtype=np.float32) - 0.2).astype(np.float32)
    thr = 0.05
    want_full = np.where(full < thr, 0.0, full).astype(np.float32)
    outs = [torch.full((per * S * S,), float("nan"), device="cuda") for _ in range(world)]
    ptrs = [o.data_ptr() for o in outs]
    for g in range(world):
        slab = torch.from_numpy(np.ascontiguousarray(full[g * per:(g + 1) * per])).cuda()
        capi.reslab_p2p(h, slab, ptrs, g, S, split, perm, thr, 0)
    torch.cuda.synchronize()
    for r in range(world):
        blk = want_full[:, :, r * per:(r + 1) * per] if split == 2 else want_full[:, r * per:(r + 1) * per, :]
        want = np.ascontiguousarray(blk.transpose(perm))
        got = outs[r].cpu().numpy().reshape(want.shape)
        assert np.array_equal(got, want), (world, split, perm, r)


@pytest.mark.parametrize("world,split,perm,batch", [(2, 2, (2, 0, 1), 8), (4, 2, (2, 0, 1), 3), (4, 2, (2, 1, 0), 8),
                                                     (2, 1, (1, 2, 0), 16)])
def test_reslab_p2p_part_streams_batches(world, split, perm, batch):
    """mpg_reslab_p2p_part: the axis change issued per finished slice batch (rows [a0, a0+count) of the old slice axis)
    gives exactly the whole-slab result; ranks emulated on one GPU as in the test above."""
    from mpgan_b200 import capi
    h = capi.default_handle(0)
    S = 96
    per = S // world
    rng = np.random.default_rng(12)
    full = (rng.random((S, S, S), dtype=np.float32) - 0.2).astype(np.float32)
    thr = 0.05
    want_full = np.where(full < thr, 0.0, full).astype(np.float32)
    outs = [torch.full((per * S * S,), float("nan"), device="cuda") for _ in range(world)]
    ptrs = [o.data_ptr() for o in outs]
    for g in range(world):
        slab = torch.from_numpy(np.ascontiguousarray(full[g * per:(g + 1) * per])).cuda()
        for a in range(0, per, batch):
            cnt = min(batch, per - a)
            capi.reslab_p2p_part(h, slab[a], ptrs, S, g * per + a, cnt, split, perm, thr, 0)
    torch.cuda.synchronize()
    for r in range(world):
        blk = want_full[:, :, r * per:(r + 1) * per] if split == 2 else want_full[:, r * per:(r + 1) * per, :]
        want = np.ascontiguousarray(blk.transpose(perm))
        got = outs[r].cpu().numpy().reshape(want.shape)
        assert np.array_equal(got, want), (world, split, perm, r)


def test_reslab_p2p_part_rejects_misaligned_rows():
    """final_perm[2] == 0 stores 4 consecutive rows of the part as one 128-bit word: a0 and count must be 4-aligned."""
    from mpgan_b200 import capi
    h = capi.default_handle(0)
    S, world = 32, 2
    outs = [torch.zeros((S // world) * S * S, device="cuda") for _ in range(world)]
    part = torch.zeros((3, S, S), device="cuda")
    with pytest.raises(capi.MpgError):
        capi.reslab_p2p_part(h, part, [o.data_ptr() for o in outs], S, 0, 3, 2, (2, 1, 0), 0.0, 0)


def test_host_frame_loop_matches_direct_calls():
    """pipeline.HostFrameLoop (overlapped H2D / D2H on copy streams, two frames in flight) returns, for every frame,
    exactly what a plain call of the pipeline returns."""
    from mpgan_b200 import pipeline as P, synth
    L = 8
    w1, w2 = P.make_weights_4x(L, 3)
    mp = P.MultiPass4x(L, w1, w2, precision="fp16", batch=8)
    frames = [synth.synthetic_volume(L, seed=s) for s in (1, 2, 3, 4, 5)]
    want = [mp(f).cpu().numpy().copy() for f in frames]
    loop = P.HostFrameLoop(mp, depth=2)
    pins = [torch.from_numpy(f).pin_memory() for f in frames]
    got = []
    slots = []
    for i, xp in enumerate(pins):
        slots.append(loop.submit(xp))
        if i >= 1:  # consume frame i-1 while frame i is in flight
            got.append(loop.result(slots[i - 1]).numpy().copy())
    got.append(loop.result(slots[-1]).numpy().copy())
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    got2 = loop.result(loop.submit(frames[0])).numpy()  # numpy input goes through the pinned staging buffer
    assert np.array_equal(got2, want[0])


@pytest.mark.parametrize("pass2", [False, True])
def test_slice_assemble_fast_path_equals_generic_kernel_at_the_benchmarked_size(pass2):
    """L = 128, upRes 4 (BASELINE config 2): the 4-channel fast path (one thread per in-plane position, all slices of the batch)
    against the generic per-pixel kernel (selected by a wider output stride), ragged batches at several slice offsets; both
    evaluate the same trilinear weights, in a different order of operations."""
    from mpgan_b200 import synth
    L, u = 128, 4
    S = L * u
    vol = torch.from_numpy(synth.synthetic_volume(L, seed=3)).cuda()
    dens = torch.rand((24, S, S), device="cuda") if pass2 else None
    geom = dict(axis_of=(2, 0, 1), zoom=(u, u, u), chans=(2, 3, 1), scale=(4.0, 4.0, 4.0)) if pass2 else \
        dict(axis_of=(0, 1, 2), zoom=(u, 1, 1), chans=(0, 1, 2, 3), scale=(1.0, 0.5, 0.5, 0.5))
    side = S if pass2 else L
    fast = capi.make_assemble_desc((L, L, L), 4, geom["axis_of"], geom["zoom"], geom["chans"], geom["scale"], out_dtype=capi.F32,
                                   out_cstride=4, dens_slice0=200 if pass2 else 0)
    slow = capi.make_assemble_desc((L, L, L), 4, geom["axis_of"], geom["zoom"], geom["chans"], geom["scale"], out_dtype=capi.F32,
                                   out_cstride=8, dens_slice0=200 if pass2 else 0)
    for s0, count in ((200, 8), (205, 3), (216, 8)) if pass2 else ((0, 8), (251, 5), (504, 8)):
        a = torch.full((count, side, side, 4), float("nan"), device="cuda")
        b = torch.full((count, side, side, 8), float("nan"), device="cuda")
        capi.slice_assemble(H(), fast, vol, dens, s0, count, a)
        capi.slice_assemble(H(), slow, vol, dens, s0, count, b)
        assert torch.isfinite(a).all() and float(b[..., 4:].abs().max()) == 0.0
        scale = float(b[..., :4].abs().max())
        assert float((a - b[..., :4]).abs().max()) <= 2e-6 * max(1.0, scale), (s0, count)
    # a slice's value does not depend on the batch it is assembled in (sharded / tiled runs stay bit-identical)
    one = torch.empty((1, side, side, 4), device="cuda")
    many = torch.empty((8, side, side, 4), device="cuda")
    s0 = 208 if pass2 else 96
    capi.slice_assemble(H(), fast, vol, dens, s0, 8, many)
    capi.slice_assemble(H(), fast, vol, dens, s0 + 5, 1, one)
    assert torch.equal(one[0], many[5])
