"""world_size-2 gloo test of the slice-sharded pipeline orchestration (host logic of parallel.py):
two ranks each own half the slices; after the all-to-all re-slabbing the concatenated result must equal
the single-rank result bit for bit (slice sharding must not change any voxel)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _permute3_cpu(src, dst, dims, perm, thr):
    v = src.reshape(-1)[: dims[0] * dims[1] * dims[2]].view(*dims).permute(*perm).contiguous()
    if thr > 0:
        v = torch.where(v < thr, torch.zeros_like(v), v)
    dst.reshape(-1)[: v.numel()].copy_(v.reshape(-1))


def _two_pass_identity(S, rank, world, group, vol1_rows, net2):
    """Orchestration of MultiPass4x.__call__ with injected 'networks' on CPU tensors."""
    sys.path.insert(0, ROOT)
    import mpgan_b200  # noqa: F401
    from mpgan_b200 import parallel as par
    s0, s1 = par.slab_range(rank, world, S)
    per = s1 - s0
    vol_a = vol1_rows[s0:s1].clone()  # pass-1 rows of this rank: [z_loc, y, x]
    vol_b = torch.empty(per, S, S)
    sa, sb = torch.empty(per, S, S), torch.empty(per, S, S)
    par.reslab(vol_a, S, world, group, _permute3_cpu, sa, sb, vol_b, (2, 0, 1), 0.0005)  # -> [x_loc, Z, Y]
    rows2 = net2(vol_b, s0)  # pass-2 rows [x_loc, Z, Y]
    out = torch.empty(per, S, S)
    par.reslab_mid(rows2, S, world, group, _permute3_cpu, sa, sb, out, (1, 2, 0), 0.0005)  # -> [z_loc, Y, X]
    return out


def _net2(vol_b, s0):
    # any slice-local function of the (x, z, y) slab and the absolute slice index
    idx = torch.arange(vol_b.shape[0], dtype=torch.float32).view(-1, 1, 1) + s0
    return vol_b * 1.5 + 0.001 * idx + torch.roll(vol_b, 1, dims=2) * 0.25


def _worker(rank, world, port, S, ref_path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vol1 = torch.from_numpy(np.load(ref_path))
    out = _two_pass_identity(S, rank, world, None, vol1, _net2)
    np.save(os.path.join(out_dir, "out_%d.npy" % rank), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_two_pass_equals_single_rank(tmp_path, world):
    S = 8
    rng = np.random.default_rng(0)
    vol1 = (rng.random((S, S, S)).astype(np.float32) * 0.002)  # straddles the 0.0005 threshold
    ref_path = str(tmp_path / "vol1.npy")
    np.save(ref_path, vol1)
    single = _two_pass_identity(S, 0, 1, None, torch.from_numpy(vol1), _net2).numpy()
    # reference semantics: threshold, transpose(2,0,1), net, transpose(1,2,0), threshold
    v = vol1.copy()
    v[v < 0.0005] = 0
    r = _net2(torch.from_numpy(np.ascontiguousarray(v.transpose(2, 0, 1))), 0).numpy().transpose(1, 2, 0).copy()
    r[r < 0.0005] = 0
    assert np.array_equal(single, r)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, S, ref_path, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(str(tmp_path / ("out_%d.npy" % r))) for r in range(world)]
    assert np.array_equal(np.concatenate(parts, axis=0), single)


def _grad_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import mpgan_b200  # noqa: F401
    from mpgan_b200 import parallel as par
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1) + 0.25 * rank  # this rank's flat gradient buffer
    par.allreduce_mean(g)
    np.save(os.path.join(out_dir, "g_%d.npy" % rank), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_training_gradient_exchange_is_the_rank_mean(tmp_path):
    """Data-parallel training step: one all-reduce over the flat gradient, averaged (parallel.allreduce_mean)."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    base = np.arange(1000, dtype=np.float32)
    want = (base * 1 + base * 2 + 0.25) / 2
    for r in range(world):
        np.testing.assert_allclose(np.load(str(tmp_path / ("g_%d.npy" % r))), want, rtol=1e-6)


def _any_worker(rank, world, port, S, ref_path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import itertools
    import mpgan_b200  # noqa: F401
    from mpgan_b200 import parallel as par
    full = torch.from_numpy(np.load(ref_path))
    s0, s1 = par.slab_range(rank, world, S)
    per = s1 - s0
    for perm in itertools.permutations(range(3)):
        out = torch.empty(per, S, S)
        sa, sb = torch.empty(per, S, S), torch.empty(per, S, S)
        par.reslab_any(full[s0:s1].clone(), S, world, None, _permute3_cpu, sa, sb, out, perm, 0.0)
        np.save(os.path.join(out_dir, "any_%d_%d%d%d.npy" % ((rank,) + perm)), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_reslab_any_covers_every_axis_permutation(tmp_path):
    """parallel.reslab_any (what the generalized sharded MultiPassOut uses for any generator chain / transposeAxis):
    for every axis permutation the ranks' outputs concatenate to full.transpose(perm); compose_perms is numpy's rule."""
    import itertools
    sys.path.insert(0, ROOT)
    import mpgan_b200  # noqa: F401
    from mpgan_b200 import parallel as par
    S, world = 8, 2
    full = np.random.default_rng(3).random((S, S, S)).astype(np.float32)
    for p, q in itertools.product(itertools.permutations(range(3)), repeat=2):
        assert np.array_equal(full.transpose(p).transpose(q), full.transpose(par.compose_perms(p, q)))
    # the closing composition of the shipped recipes (GAN/multipassGAN-out.py:521,587-590)
    assert par.compose_perms((1, 2, 0), (2, 0, 1), (2, 1, 0)) == (2, 1, 0)
    assert par.compose_perms((0, 1, 2), (2, 0, 1), (2, 1, 0)) == (1, 0, 2)
    assert par.compose_perms((2, 1, 0), (2, 1, 0)) == (0, 1, 2)
    ref_path = str(tmp_path / "full.npy")
    np.save(ref_path, full)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_any_worker, args=(world, port, S, ref_path, str(tmp_path)), nprocs=world, join=True)
    for perm in itertools.permutations(range(3)):
        parts = [np.load(str(tmp_path / ("any_%d_%d%d%d.npy" % ((r,) + perm)))) for r in range(world)]
        assert np.array_equal(np.concatenate(parts, axis=0), full.transpose(perm)), perm
