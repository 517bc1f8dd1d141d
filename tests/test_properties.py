"""Property tests (hypothesis): the data-movement steps of the path are exact permutations /
exact inverses for arbitrary sizes -- the pack -> all-to-all -> unpack axis change run by G rank threads, the overlapped
tile cut/stitch, the .uni codecs, the checkpoint primitives, the tile sampler's bounds."""
import random
import threading

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import mpgan_b200  # noqa: F401
from mpgan_b200 import io_pipeline, parallel as par, tfckpt, tilesampler, uni
from oracle import tiles as otiles

FAST = settings(max_examples=20, deadline=None)


def _permute3(src, dst, dims, perm, thr):
    v = src.reshape(dims).permute(*perm)
    if thr > 0:
        v = torch.where(v < thr, torch.zeros_like(v), v)
    dst.view(-1).copy_(v.reshape(-1))


class _FakeExchange:
    """all_to_all_single between G threads of one process: chunk h of rank g's send buffer becomes chunk g of rank h's
    receive buffer."""

    def __init__(self, world):
        self.world, self.bar, self.send = world, threading.Barrier(world), {}
        self.local = threading.local()

    def __call__(self, recv, send, group):
        g = self.local.rank
        self.send[g] = send.view(-1)
        self.bar.wait()
        n = send.numel() // self.world
        for h in range(self.world):
            recv.view(-1)[h * n:(h + 1) * n] = self.send[h][g * n:(g + 1) * n]
        self.bar.wait()


@FAST
@given(world=st.sampled_from([1, 2, 3, 4]), per=st.integers(1, 5), which=st.sampled_from([0, 1, 2]), thr=st.sampled_from([0.0, 0.3]),
       seed=st.integers(0, 1000))
def test_reslab_is_the_exact_axis_change(world, per, which, thr, seed):
    """parallel.reslab / reslab_mid executed by `world` rank threads == permute of the full volume restricted to the
    rank's new slab, for the three (split, permutation) pairs the pipelines use."""
    S = world * per
    rng = np.random.default_rng(seed)
    full = rng.random((S, S, S), dtype=np.float32)
    fn, perm = [(par.reslab, (2, 0, 1)), (par.reslab, (2, 1, 0)), (par.reslab_mid, (1, 2, 0))][which]
    ex = _FakeExchange(world)
    outs = [None] * world
    errs = []

    def rank_main(g):
        try:
            ex.local.rank = g
            slab = torch.from_numpy(np.ascontiguousarray(full[g * per:(g + 1) * per]))
            a, b, out = (torch.empty(per * S * S) for _ in range(3))
            fn(slab, S, world, None, _permute3, a, b, out, perm, thr, all_to_all=ex)
            outs[g] = out.numpy().copy()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            ex.bar.abort()

    threads = [threading.Thread(target=rank_main, args=(g,)) for g in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    want_full = np.where(full < thr, 0.0, full).astype(np.float32) if thr > 0 else full
    for h in range(world):
        blk = want_full[:, :, h * per:(h + 1) * per] if fn is par.reslab else want_full[:, h * per:(h + 1) * per, :]
        assert np.array_equal(outs[h].reshape(blk.transpose(perm).shape), blk.transpose(perm))


@FAST
@given(n=st.integers(1, 3), ty=st.integers(1, 4), tx=st.integers(1, 4), core=st.integers(1, 6), b=st.integers(0, 3), c=st.integers(1, 3),
       seed=st.integers(0, 1000))
def test_overlap_cut_then_stitch_is_identity(n, ty, tx, core, b, c, seed):
    tile = core + 2 * b
    rng = np.random.default_rng(seed)
    x = rng.random((n, ty * core + 2 * b, tx * core + 2 * b, c), dtype=np.float32)
    tiles = otiles.cut_overlap(x, tile, b) if ty == tx else None
    if tiles is None:  # cut_overlap assumes one tile size per axis pair; build the general grid directly
        tiles = np.stack([x[i, y * core:y * core + tile, xx * core:xx * core + tile] for i in range(n) for y in range(ty)
                          for xx in range(tx)])
    assert np.array_equal(otiles.stitch_overlap(tiles, n, ty, tx, b), x)


@FAST
@given(dz=st.integers(1, 5), dy=st.integers(1, 6), dx=st.integers(1, 7), vec=st.booleans(), chunk=st.integers(64, 4096),
       seed=st.integers(0, 1000))
def test_uni_writers_roundtrip(tmp_path_factory, dz, dy, dx, vec, chunk, seed):
    d = tmp_path_factory.mktemp("uni")
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((dz, dy, dx, 3 if vec else 1)).astype(np.float32)
    head = uni.make_header((dz, dy, dx), 2 if vec else 1, timestamp=seed)
    uni.write_uni(str(d / "a.uni"), head, data)
    io_pipeline.write_uni_parallel(str(d / "b.uni"), head, data, threads=2, chunk_bytes=chunk)
    for name in ("a.uni", "b.uni"):
        h, v = uni.read_uni(str(d / name))
        assert h == head and np.array_equal(v, data)


@FAST
@given(v=st.integers(0, 2 ** 64 - 1), data=st.binary(max_size=300), split=st.integers(0, 300))
def test_checkpoint_primitives(v, data, split):
    enc = tfckpt.put_varint(v)
    assert tfckpt.get_varint(enc + b"\x00", 0) == (v, len(enc))
    crc = tfckpt.crc32c(data)
    assert tfckpt.unmask_crc(tfckpt.mask_crc(crc)) == crc
    k = min(split, len(data))
    assert tfckpt.crc32c(data[k:], tfckpt.crc32c(data[:k])) == crc


@FAST
@given(T=st.integers(1, 6), extra=st.integers(0, 6), u=st.sampled_from([1, 2, 4]), n=st.integers(1, 4), seed=st.integers(0, 1000))
def test_tile_sampler_picks_stay_inside_and_pair_up(T, extra, u, n, seed):
    L = T + extra
    rng = np.random.default_rng(seed)
    low = rng.random((n + 1, 1, L, L, 4), dtype=np.float32)
    high = rng.random((n + 1, 1, L * u, L * u, 1), dtype=np.float32)
    s = tilesampler.TileSampler(T, u, densityMinimum=0.3, partTrain=0.9, partTest=0.1, rng=random.Random(seed))
    s.add_data(low, high)
    if s.set_borders[0] < 1:
        return
    picks = s.select_offsets(5)
    assert all(0 <= f < s.set_borders[0] and 0 <= oy <= L - T and 0 <= ox <= L - T for f, oy, ox, _ in picks)
    lo, hi = s.gather(picks)
    for i, (f, oy, ox, _) in enumerate(picks):
        assert np.array_equal(lo[i, 0].numpy(), low[f, 0, oy:oy + T, ox:ox + T])
        assert np.array_equal(hi[i, 0].numpy(), high[f, 0, oy * u:(oy + T) * u, ox * u:(ox + T) * u])
