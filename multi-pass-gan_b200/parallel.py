"""Slice sharding of the multi-pass pipeline across the GPUs of one box.

Within a pass every slice is an independent 2-D inference (GAN/multipassGAN-out.py:443-447), so rank g
owns the contiguous slice range [g*S/G, (g+1)*S/G) of the pass's slice axis.  The only communication is
the axis change between passes (the `np.array(rows).reshape(S,S,S).transpose(...)` of
GAN/multipassGAN-out.py:459,521 / GAN/multipassGAN-4x.py:1142): a slab along the old slice axis must
become a slab along the new one = ONE all-to-all of S^3/G^2-element blocks per rank pair.  On the GPUs this
is a single kernel: `PeerSlab.exchange` transposes the slab and stores every element straight into the output
slab of the owning rank (symmetric memory mapped over NVLink 5 / NVSwitch, capi.reslab_p2p).  `reslab` /
`reslab_mid` are the pack -> all_to_all_single -> unpack formulation of the same exchange: NCCL on the GPUs
when MPG_EXCHANGE=nccl, "gloo" on CPU tensors in the host-logic tests.

`permute3(src, dst, dims, perm, threshold)` is the 3-D axis permutation primitive: on the GPU it is
capi.transpose3d (mpg_transpose3d), in the CPU tests a torch.permute.
"""
import torch
import torch.distributed as dist


def slab_range(rank, world, n):
    """Contiguous slice range owned by `rank`; requires world | n (the reference drops ragged batches)."""
    if n % world:
        raise ValueError("slice count %d is not divisible by the number of ranks %d" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def _all_to_all(recv, send, group):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    else:
        recv.view(-1).copy_(send.view(-1))


def reslab(slab, S, world, group, permute3, scratch_a, scratch_b, out, final_perm, threshold=0.0, all_to_all=None):
    """Turn this rank's slab along axis 0 into its slab along (old) axis 2, permuted by `final_perm`.

    slab      [S/G, S, S]   (a_loc, b, c)  rows of the finished pass, a = old slice axis
    out       the rank's part of the re-sliced volume:  permute(recv[(A), b, c_loc], final_perm)
              where recv is the full-A extent restricted to this rank's c range.
    scratch_* two buffers of S^3/G elements.
    all_to_all(recv, send, group): the collective (default: torch.distributed all_to_all_single); tests inject an
              in-process exchange between rank threads.
    Steps: pack [a_loc*b, G, c_loc] -> [G, a_loc*b, c_loc]; all-to-all; the received chunks, ordered by
    source rank, ARE [A, b, c_loc]; one final permutation (with the optional threshold fused).
    """
    per = S // world
    if world == 1:
        permute3(slab, out, (S, S, S), final_perm, threshold)
        return out
    permute3(slab, scratch_a, (per * S, world, per), (1, 0, 2), 0.0)
    (all_to_all or _all_to_all)(scratch_b, scratch_a, group)
    permute3(scratch_b, out, (S, S, per), final_perm, threshold)
    return out


def reslab_mid(slab, S, world, group, permute3, scratch_a, scratch_b, out, final_perm, threshold=0.0, all_to_all=None):
    """Same as `reslab` but the NEW slab axis is the old axis 1 (b):  [a_loc, B, c] -> perm([A, b_loc, c]).

    pack [a_loc, G, b_loc*c] -> [G, a_loc, b_loc*c]; all-to-all -> [A, b_loc, c]; final permutation.
    """
    per = S // world
    if world == 1:
        permute3(slab, out, (S, S, S), final_perm, threshold)
        return out
    permute3(slab, scratch_a, (per, world, per * S), (1, 0, 2), 0.0)
    (all_to_all or _all_to_all)(scratch_b, scratch_a, group)
    permute3(scratch_b, out, (S, per, S), final_perm, threshold)
    return out


def compose_perms(*perms):
    """x.transpose(p1).transpose(p2)... as ONE numpy-style axis permutation."""
    cur = (0, 1, 2)
    for p in perms:
        cur = tuple(cur[i] for i in p)
    return cur


def reslab_any(slab, S, world, group, permute3, scratch_a, scratch_b, out, perm, threshold=0.0, all_to_all=None):
    """out = this rank's axis-0 slab of full.transpose(perm), `slab` being its axis-0 slab of `full` [S,S,S].
    perm[0] == 0: the slab axis stays (local permutation of [S/G,S,S]); 2 -> `reslab`; 1 -> `reslab_mid`."""
    per = S // world
    if perm[0] == 0:
        permute3(slab, out, (per, S, S), perm, threshold)
        return out
    fn = reslab if perm[0] == 2 else reslab_mid
    return fn(slab, S, world, group, permute3, scratch_a, scratch_b, out, perm, threshold, all_to_all)


class PeerSlab:
    """An output slab allocated in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations mapped
    into every rank of the group over NVLink / NVSwitch), so that the axis change between passes is a kernel that
    stores straight into the slab of the rank owning each element (capi.reslab_p2p[_part]) instead of pack ->
    all-to-all -> unpack.

    Protocol of the pipelines (ONE cross-rank barrier per pass boundary):
      * `push_part(..)` after every finished slice batch: that batch's rows are transposed and stored into their
        owners' slabs while the rest of the pass is still running (no barrier before it: see below)
      * `landed()` at the pass boundary: device-side barrier -- every rank's stores are visible before anyone reads
      * the two axis changes of a frame target two DIFFERENT symmetric slabs (`PeerSlab` instances), so nobody
        stores into a slab a peer may still be reading in the current pass; the slab written by the first exchange of
        frame f+1 was last read before the closing barrier of frame f.
    `exchange(..)` is the whole-slab form (barrier before only when `wait_readers`)."""

    def __init__(self, shape, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.tensor = symm_mem.empty(tuple(shape), dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.tensor, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)

    def push_part(self, capi, handle, part, S, a0, count, split_axis, final_perm, threshold=0.0):
        st = torch.cuda.current_stream(self.tensor.device).cuda_stream
        capi.reslab_p2p_part(handle, part, self.ptrs, S, a0, count, split_axis, final_perm, threshold, st)

    def landed(self, channel=0):
        self.hdl.barrier(channel=channel)
        return self.tensor

    def exchange(self, capi, handle, slab, S, split_axis, final_perm, threshold=0.0, wait_readers=True):
        st = torch.cuda.current_stream(self.tensor.device).cuda_stream
        if wait_readers:
            self.hdl.barrier(channel=0)
        capi.reslab_p2p(handle, slab, self.ptrs, self.rank, S, split_axis, final_perm, threshold, st)
        self.hdl.barrier(channel=1)
        return self.tensor


def make_peer_slab(shape, device, S, world, group=None):
    """PeerSlab when the fused exchange is usable, else None (-> NCCL all-to-all path). A failing symmetric-memory
    rendezvous (e.g. no P2P mapping between the visible GPUs) is reported on stderr and ALL ranks fall back together:
    the rendezvous is collective, and the outcome is agreed on with an all-reduce before anyone proceeds."""
    if not p2p_usable(S, world):
        return None
    import sys
    slab, ok = None, 1
    try:
        slab = PeerSlab(shape, device, group)
    except Exception as e:  # noqa: BLE001
        ok = 0
        sys.stderr.write("mpgan_b200: symmetric-memory exchange unavailable (%s: %s); using the NCCL all-to-all path\n"
                         % (type(e).__name__, e))
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return slab if int(flag.item()) == 1 else None


def p2p_usable(S, world):
    """The fused peer-store exchange needs CUDA, an initialised NCCL group and S/G % 4 == 0 (128-bit stores)."""
    import os
    if os.environ.get("MPG_EXCHANGE", "p2p") != "p2p":
        return False
    return (world > 1 and torch.cuda.is_available() and dist.is_available() and dist.is_initialized()
            and dist.get_backend() == "nccl" and (S // world) % 4 == 0)


def allreduce_mean(flat, group=None):
    """Data-parallel gradient exchange of the training step: ONE all-reduce over the flat gradient buffer of an
    optimizer, averaged over the ranks (each rank draws its own tile batch; BN batch statistics stay per rank)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / dist.get_world_size(group))
    return flat


def init_from_env(device_index=None):
    """One process per GPU under torchrun: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if torch.cuda.is_available():
            torch.cuda.set_device(local if device_index is None else device_index)
            dist.init_process_group("nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local if device_index is None else device_index))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, local, world
