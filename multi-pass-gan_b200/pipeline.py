"""Device-resident multi-pass volume pipelines (the per-frame hot loop of the reference).

  MultiPass4x   the shipped 4x recipe, GAN/example_run_output.py:6,8 = two runs of
                GAN/multipassGAN-4x.py:1090-1169 (pass 1 `upsamplingMode 2`, pass 2 `upsamplingMode 1`)
  MultiPassOut  generate3DUniForNewNetwork of GAN/multipassGAN-out.py:390-618 (1-3 chained generators)

In the reference every slice batch crosses the host boundary twice (feed + fetch of `sess.run`) and the
volume lives in host numpy memory between passes; here the low-res fields are uploaded once, slice
batches are assembled on the device (mpg_slice_assemble), each network writes its rows straight into the
pass volume, and the axis changes between passes are mpg_transpose3d launches.  Slices of one pass are
independent, so a rank may own any contiguous slice range (`slice_range`); the exchange between passes
is then an all-to-all (parallel.py).
"""
import numpy as np
import torch

from . import capi, engine
from . import parallel as par
from . import graph as G
from . import networks as N
from . import weights as W

THRESHOLD = 0.0005  # GAN/multipassGAN-out.py:614, GAN/multipassGAN-4x.py:1156

# GAN/multipassGAN-out.py:398-421 / :464-485 / :526-547 as (axis_of, channel permutation) per transposeAxis:
# output index (slice,row,col) -> source axis of the [Z,Y,X,C] field volume; the zoomed axis is axis_of[0].
_PASS_GEOM = {
    1: {0: ((0, 1, 2), (0, 1, 2, 3)), 1: ((1, 0, 2), (0, 1, 3, 2)), 2: ((2, 1, 0), (0, 3, 2, 1)),
        3: ((2, 0, 1), (0, 2, 3, 1))},
    2: {3: ((1, 0, 2), (0, 1, 3, 2)), 0: ((2, 1, 0), (0, 3, 2, 1)), 1: ((2, 0, 1), (0, 2, 3, 1)),
        2: ((0, 1, 2), (0, 1, 2, 3))},
    # pass 3: transposeAxis 2 indexes channel 13 and 3 reshapes a mis-shaped array in the reference (dead branches)
    3: {0: ((1, 0, 2), (0, 1, 3, 2)), 1: ((0, 1, 2), (0, 1, 2, 3))},
}


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def _pick_batch(n, pref):
    """Largest batch <= pref dividing n (the reference drops the remainder, App. D.4)."""
    b = min(pref, n)
    while n % b:
        b -= 1
    return b


class _PassNet:
    """One generator compiled for a fixed slice batch, fed by the slice assembler."""

    def __init__(self, handle, build_fn, weights, batch, precision, range_check=False):
        G.reset_default_graph()
        self.out_t = build_fn()
        self.graph = G.get_default_graph()
        self.net = engine.CompiledNet(self.out_t, weights, batch, precision=precision, handle=handle, range_check=range_check)
        self.batch = batch


class _TiledPassNet:
    """A generator applied to whole [H,H] slices through overlapping [T,T] tiles (the reference's
    TileCreator.createTiles / concatTiles, tools_wscale/tilecreator_t.py:403-450,886-918, used the B200 way):
    mpg_tiles_cut (stride = core, tile = core + 2*halo) -> ONE network launch over batch*nt*nt tiles ->
    mpg_tiles_stitch_overlap (crop `halo`, keep the outer band of frame-edge tiles). With halo >= the receptive-field
    radius of the generator (16 output pixels for gen_resnet, App. A.6) the result is bit-identical to the whole-slice
    apply, so slices that exceed HBM can be processed at any tile size."""

    def __init__(self, handle, device, build_fn, weights, batch, precision, H, core, halo, cin, scale):
        if (H - 2 * halo) % core or core <= 0:
            raise ValueError("tiled apply: (H - 2*halo) = %d is not a multiple of the tile core %d" % (H - 2 * halo, core))
        self.h, self.H, self.core, self.halo, self.cin, self.scale = handle, H, core, halo, cin, scale
        self.T = core + 2 * halo
        self.nt = (H - 2 * halo) // core
        self.batch = batch
        nb = batch * self.nt * self.nt
        G.reset_default_graph()
        self.out_t = build_fn(self.T)
        self.net = engine.CompiledNet(self.out_t, weights, nb, precision=precision, handle=handle)
        f32 = dict(dtype=torch.float32, device=device)
        self.tin = torch.empty((nb, self.T, self.T, cin), **f32)
        self.tout = torch.empty((nb, self.T * scale, self.T * scale), **f32)

    def run(self, feeds, out, stream):
        x = feeds["x"]
        capi.tiles_cut(self.h, x, self.tin, self.batch, self.H, self.H, self.cin, 4, self.T, self.T, self.core, self.core, 0,
                       stream)
        self.net.run({"x": self.tin}, out=self.tout, stream=stream)
        capi.tiles_stitch_overlap(self.h, self.tout, out, self.batch, self.nt, self.nt, self.T * self.scale,
                                  self.T * self.scale, 1, 4, self.halo * self.scale, stream)

    @property
    def flops(self):
        return self.net.flops

    @property
    def launches(self):
        return self.net.launches + 2


class _TiledHolder:
    def __init__(self, net):
        self.net = net


class MultiPass4x:
    """Two-pass 4x super-resolution of one frame: [L,L,L,4] -> [4L,4L,4L] (z,y,x), fp32.

    With world > 1 the volume is sharded by slice: the rank computes z-slab [S/G,S,S] in pass 1, the
    all-to-all turns z-slabs into x-slabs for pass 2, and a second one returns canonical z-slabs, so
    `__call__` returns this rank's [S/G, S, S] part of the output (rank-major == z order)."""

    def __init__(self, L, weights_pass1, weights_pass2, upRes=4, precision="fp16", batch=8, velScale=1.0,
                 batch_norm=True, device=0, threshold=THRESHOLD, rank=0, world=1, group=None, tile=None, range_check=False):
        """tile: None = whole slices per launch (the reference's behaviour), or (core1, core2): pass 1 is applied to
        overlapping low-res tiles of core1 + 2*4 pixels, pass 2 to high-res tiles of core2 + 2*16 pixels
        (_TiledPassNet; (L - 8) % core1 == 0, (S - 32) % core2 == 0, core2 % upRes == 0) -- same result, bounded memory.
        range_check: validation mode -- every 16-bit layer output is scanned for saturated stores; a call raises
        capi.MpgRangeError instead of returning clipped data (engine.CompiledNet)."""
        self.L, self.u, self.S = int(L), int(upRes), int(L) * int(upRes)
        self.h = capi.default_handle(device)
        self.device = torch.device("cuda", device)
        self.precision = precision
        self.velScale = float(velScale)
        self.threshold = float(threshold)
        self.rank, self.world, self.group = int(rank), int(world), group
        L, S, u = self.L, self.S, self.u
        self.s0, self.s1 = par.slab_range(self.rank, self.world, S)
        self.S_loc = self.s1 - self.s0
        self.batch = _pick_batch(self.S_loc, batch)
        cfg1 = N.config_4x(L, upRes=u, upsampling_mode=2, batch_norm=batch_norm)
        cfg2 = N.config_4x(L, upRes=u, upsampling_mode=1, batch_norm=batch_norm)
        # either weight set may be None: the reference runs the two passes as two processes that hand the volume over as
        # a .uni file (GAN/example_run_output.py:6,8); pass1_only / pass2_only are those two runs (cli_4x.py)
        if tile is not None and (weights_pass1 is None or weights_pass2 is None):
            raise ValueError("tiled apply needs both generators")
        self.p1 = _PassNet(self.h, lambda: N.gen_resnet(G.placeholder([None, L * L * 4], "x"), cfg1), weights_pass1,
                           self.batch, precision, range_check) if weights_pass1 is not None else None
        self.p2 = _PassNet(self.h, lambda: N.gen_resnet(G.placeholder([None, S * S * 4], "x"), cfg2), weights_pass2,
                           self.batch, precision, range_check) if (tile is None and weights_pass2 is not None) else None
        if tile is not None:
            core1, core2 = int(tile[0]), int(tile[1])
            halo_hi = 16  # receptive-field radius of gen_resnet: 8 convs of k=5 (App. A.6)
            if halo_hi % u or core2 % u:
                raise ValueError("tiled apply: upRes must divide the halo (16) and the pass-2 tile core")
            self.p1.net.close()

            def build1(T):
                return N.gen_resnet(G.placeholder([None, T * T * 4], "x"),
                                    N.config_4x(T, upRes=u, upsampling_mode=2, batch_norm=batch_norm))

            def build2(T):
                return N.gen_resnet(G.placeholder([None, T * T * 4], "x"),
                                    N.config_4x(T // u, upRes=u, upsampling_mode=1, batch_norm=batch_norm))

            self.p1 = _TiledHolder(_TiledPassNet(self.h, self.device, build1, weights_pass1, self.batch, precision, L, core1,
                                                 halo_hi // u, 4, u))
            self.p2 = _TiledHolder(_TiledPassNet(self.h, self.device, build2, weights_pass2, self.batch, precision, S, core2,
                                                 halo_hi, 4, 1))
        vs = self.velScale
        # pass 1: zoom(x,[u,1,1,1]) (GAN/multipassGAN-4x.py:1103); velocities * velScale (:283)
        self.asm1 = capi.make_assemble_desc((L, L, L), 4, (0, 1, 2), (u, 1, 1), (0, 1, 2, 3), (1.0, vs, vs, vs),
                                            out_dtype=capi.F32, out_cstride=4)
        # pass 2: concat(x_2, zoom(vel*u,[u,u,u,1])).transpose(0,3,1,2,4) + swaps 2<->3, 3<->1 (:1113-1119):
        # slices along x of (z,y) planes, channels (d, vy, vz, vx); velScale hits vy,vz only (App. D.10)
        self.asm2 = capi.make_assemble_desc((L, L, L), 4, (2, 0, 1), (u, u, u), (2, 3, 1),
                                            (u * vs, u * vs, float(u)), out_dtype=capi.F32, out_cstride=4,
                                            dens_slice0=self.s0)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.in1 = torch.empty((self.batch, L, L, 4), **f32)
        self.in2 = torch.empty((self.batch, S, S, 4), **f32)
        self.vol_a = torch.empty((self.S_loc, S, S), **f32)
        # vol_b: the first-pass volume re-sliced along x (input of pass 2); vol_c: the finished frame. Two buffers, so the
        # second axis change never stores into memory a peer (or a pending D2H copy) may still be reading (parallel.PeerSlab)
        self.peer = par.make_peer_slab((self.S_loc, S, S), self.device, S, self.world, group)
        self.peer_c = par.make_peer_slab((self.S_loc, S, S), self.device, S, self.world, group) if self.peer else None
        if self.peer is not None and self.peer_c is None:
            self.peer = None
        self.vol_b = self.peer.tensor if self.peer else torch.empty((self.S_loc, S, S), **f32)
        self.vol_c = self.peer_c.tensor if self.peer else torch.empty((self.S_loc, S, S), **f32)
        if self.world > 1 and self.peer is None:  # NCCL all-to-all path: pack / receive staging
            self.scr_a = torch.empty((self.S_loc, S, S), **f32)
            self.scr_b = torch.empty((self.S_loc, S, S), **f32)
        else:
            self.scr_a = self.scr_b = None
        nb = self.S_loc // self.batch
        have = [p_ for p_ in (self.p1, self.p2) if p_ is not None]
        self.flops = sum(p_.net.flops for p_ in have) / self.batch * self.S_loc  # this rank's share
        self.launches_per_frame = nb * (sum(p_.net.launches for p_ in have) + 2) + (2 if world == 1 else (nb + 1 if self.peer else 4))
        self.events = None

    def _permute3(self, src, dst, dims, perm, thr):
        capi.transpose3d(self.h, src, dst, dims, perm, thr, torch.cuda.current_stream(self.device).cuda_stream)

    def upload(self, x):
        return _dev_f32(x, self.device)

    def __call__(self, x, record=False, output_free=None):
        """x: [L,L,L,4] float32 (numpy or device tensor, replicated on every rank).
        Returns the device tensor [S/G,S,S] (z,y,x): this rank's z-slab of the output (`vol_c`, overwritten by the next
        call). output_free: optional CUDA event recorded by whoever still reads the previous call's result (e.g. a D2H
        copy on another stream); nothing stores into it before that event (HostFrameLoop)."""
        S, B = self.S, self.batch
        cur = torch.cuda.current_stream(self.device)
        st = cur.cuda_stream
        vol = _dev_f32(x, self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        if ev:
            ev[0].record()
        # ---- pass 1: xy slices along (interpolated) z; rows land in vol_a[z - s0] = [Zu_loc, Yu, Xu]. The .uni
        #      hand-over between the two processes of the reference = threshold (:1155-1157) + axis change to slices
        #      along x, [Zu,Yu,Xu] -> [Xu_loc, Zu, Yu]: with peers every finished batch is pushed to its owners at once
        for s in range(self.s0, self.s1, B):
            capi.slice_assemble(self.h, self.asm1, vol, None, s, B, self.in1, st)
            self.p1.net.run({"x": self.in1}, out=self.vol_a[s - self.s0], stream=st)
            if self.peer:
                self.peer.push_part(capi, self.h, self.vol_a[s - self.s0], S, s, B, 2, (2, 0, 1), self.threshold)
        if ev:
            ev[1].record()
        if output_free is not None:
            cur.wait_event(output_free)
        if self.peer:
            self.peer.landed(channel=0)  # the ONE barrier of this pass boundary: all ranks' stores are visible
        elif self.world > 1:
            par.reslab(self.vol_a, S, self.world, self.group, self._permute3, self.scr_a, self.scr_b, self.vol_b,
                       (2, 0, 1), self.threshold)
        else:
            self._permute3(self.vol_a, self.vol_b, (S, S, S), (2, 0, 1), self.threshold)
        if ev:
            ev[2].record()
        for s in range(self.s0, self.s1, B):
            capi.slice_assemble(self.h, self.asm2, vol, self.vol_b, s, B, self.in2, st)
            self.p2.net.run({"x": self.in2}, out=self.vol_a[s - self.s0], stream=st)
        if ev:
            ev[3].record()
        # rows [Xu_loc, Zu, Yu] -> .transpose(1,2,0) -> [Zu_loc, Yu, Xu] (:1142), threshold (:1155-1157); the target is
        # vol_c, which no rank reads during pass 2, so no barrier is needed BEFORE the stores
        if self.peer:
            self.peer_c.exchange(capi, self.h, self.vol_a, S, 1, (1, 2, 0), self.threshold, wait_readers=False)
        elif self.world > 1:
            par.reslab_mid(self.vol_a, S, self.world, self.group, self._permute3, self.scr_a, self.scr_b, self.vol_c,
                           (1, 2, 0), self.threshold)
        else:
            self._permute3(self.vol_a, self.vol_c, (S, S, S), (1, 2, 0), self.threshold)
        if ev:
            ev[4].record()
            self.events = ev
        return self.vol_c

    def pass1_only(self, x):
        """First-pass volume [Zu_loc,Yu,Xu] after the threshold (what pass 2 reads from density_low_2x2_*.uni)."""
        S, B = self.S, self.batch
        st = torch.cuda.current_stream(self.device).cuda_stream
        vol = _dev_f32(x, self.device)
        for s in range(self.s0, self.s1, B):
            capi.slice_assemble(self.h, self.asm1, vol, None, s, B, self.in1, st)
            self.p1.net.run({"x": self.in1}, out=self.vol_a[s - self.s0], stream=st)
        capi.threshold(self.h, self.vol_a, self.S_loc * S * S, self.threshold, st)
        return self.vol_a

    def pass2_only(self, x, dens):
        """Second run of the reference recipe (`upsamplingMode 1 upsampledData 1`, GAN/multipassGAN-4x.py:1094-1146):
        x [L,L,L,4] low-res fields (only the velocities are used), dens [Zu,Yu,Xu] the first-pass volume as read back from
        density_low_2x2_%04d.uni. Returns the device volume [Zu,Yu,Xu] after the threshold. Single GPU."""
        if self.world != 1:
            raise NotImplementedError("pass2_only is the single-process hand-over of the reference; sharded runs use __call__")
        S, B = self.S, self.batch
        st = torch.cuda.current_stream(self.device).cuda_stream
        vol = _dev_f32(x, self.device)
        d = _dev_f32(dens, self.device).reshape(S, S, S)
        self._permute3(d, self.vol_b, (S, S, S), (2, 0, 1), 0.0)  # [Zu,Yu,Xu] -> slices along x of (z,y) planes
        for s in range(0, S, B):
            capi.slice_assemble(self.h, self.asm2, vol, self.vol_b, s, B, self.in2, st)
            self.p2.net.run({"x": self.in2}, out=self.vol_a[s], stream=st)
        self._permute3(self.vol_a, self.vol_c, (S, S, S), (1, 2, 0), self.threshold)
        return self.vol_c

    def pass_times_ms(self):
        e = self.events
        return dict(pass1=e[0].elapsed_time(e[1]), exchange1=e[1].elapsed_time(e[2]), pass2=e[2].elapsed_time(e[3]),
                    exchange2=e[3].elapsed_time(e[4]))


class HostFrameLoop:
    """The frame loop of the reference scripts (GAN/multipassGAN-out.py:629-632, GAN/multipassGAN-4x.py:1634-1646) with
    HOST buffers, the B200 way: the low-res fields of frame i+1 are uploaded and the finished volume of frame i-1 is
    downloaded on two copy streams while the networks of frame i run (the reference does feed -> run -> fetch serially
    for every slice batch). `mp` is a MultiPass4x / MultiPassOut; buffers are pre-allocated and pinned, `depth` frames
    may be in flight.

        loop = HostFrameLoop(mp)
        for x in frames: k = loop.submit(x)          # x: pinned [L,L,L,4] float32 tensor (or numpy: staged through one)
        vol = loop.result(k)                          # pinned [S/G,S,S] host tensor, valid until slot k is reused
    """

    def __init__(self, mp, depth=2):
        self.mp, self.depth, self.device = mp, int(depth), mp.device
        L, S = mp.L, mp.S
        self.up, self.down = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.x_dev = [torch.empty((L, L, L, 4), dtype=torch.float32, device=self.device) for _ in range(self.depth)]
        self.x_pin = [torch.empty((L, L, L, 4), dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        self.out_pin = [torch.empty((mp.S_loc, S, S), dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        self.ev_up = [torch.cuda.Event() for _ in range(self.depth)]       # upload of slot k finished
        self.ev_used = [torch.cuda.Event() for _ in range(self.depth)]     # networks finished reading x_dev[k]
        self.ev_down = [torch.cuda.Event() for _ in range(self.depth)]     # download into out_pin[k] finished
        self.count = 0
        self.h2d_bytes = L * L * L * 4 * 4
        self.d2h_bytes = mp.S_loc * S * S * 4

    def submit(self, x):
        k = self.count % self.depth
        comp = torch.cuda.current_stream(self.device)
        if not (isinstance(x, torch.Tensor) and x.is_pinned()):
            if self.count >= self.depth:
                self.ev_up[k].synchronize()  # the staging buffer is free once its previous upload has finished
            self.x_pin[k].copy_(torch.as_tensor(x, dtype=torch.float32))
            x = self.x_pin[k]
        if self.count >= self.depth:
            self.up.wait_event(self.ev_used[k])
        with torch.cuda.stream(self.up):
            self.x_dev[k].copy_(x, non_blocking=True)
            self.ev_up[k].record(self.up)
        comp.wait_event(self.ev_up[k])
        prev = self.ev_down[(self.count - 1) % self.depth] if self.count > 0 else None
        res = self.mp(self.x_dev[k], output_free=prev)
        self.ev_used[k].record(comp)
        self.down.wait_event(self.ev_used[k])
        with torch.cuda.stream(self.down):
            self.out_pin[k].copy_(res, non_blocking=True)
            self.ev_down[k].record(self.down)
        self.count += 1
        return k

    def result(self, k):
        self.ev_down[k].synchronize()
        return self.out_pin[k]

    def drain(self):
        """Make the current stream wait for every pending download (so an event recorded next covers the whole loop)."""
        comp = torch.cuda.current_stream(self.device)
        for e in self.ev_down[:min(self.count, self.depth)]:
            comp.wait_event(e)


def make_weights_4x(L, seed, upRes=4, batch_norm=True, randomize_bn=False):
    """Random-init weights for the two 4x generators (separate variables per pass, same names)."""
    out = []
    for p, mode in ((1, 2), (2, 1)):
        G.reset_default_graph()
        cfg = N.config_4x(L, upRes=upRes, upsampling_mode=mode, batch_norm=batch_norm)
        n_in = L * L * 4 if mode == 2 else (L * upRes) ** 2 * 4
        N.gen_resnet(G.placeholder([None, n_in], "x"), cfg)
        w = W.init_graph_variables(G.get_default_graph(), seed * 10 + p)
        if randomize_bn:
            w = W.randomize_bn_stats(w, seed * 10 + p)
        out.append(w)
    return out


class NetSpec:
    """Per-network flags of GAN/multipassGAN-out.py (use_res_netN, add_adj_idcsN, startFmsN, maxFmsN, filterSizeN)."""

    def __init__(self, use_res_net=False, add_adj_idcs=False, startFms=512, maxFms=256, filterSize=3,
                 first_nn_arch=False):
        self.use_res_net, self.add_adj_idcs = bool(use_res_net), bool(add_adj_idcs)
        self.startFms, self.maxFms, self.filterSize = int(startFms), int(maxFms), int(filterSize)
        self.first_nn_arch = bool(first_nn_arch)


SHIPPED_8X = {  # GAN/example_run_output.py:18-47
    1: NetSpec(use_res_net=True, add_adj_idcs=True, startFms=256, maxFms=256, filterSize=3, first_nn_arch=True),
    2: NetSpec(use_res_net=True, add_adj_idcs=False, startFms=192, maxFms=192, filterSize=5),
    3: NetSpec(use_res_net=False, add_adj_idcs=False, startFms=192, maxFms=96, filterSize=5),
}


def build_out_graph(idx, spec, cfg):
    """Sampler wiring of GAN/multipassGAN-out.py:349-365 for generator `idx` (1-based)."""
    L, S, C = cfg.tileSizeLow, cfg.tileSizeHigh, cfg.n_inputChannels
    cu = N.log2i(cfg.upRes)
    with G.variable_scope("gen_%d" % idx):
        if idx == 1:
            cin = C + (2 if spec.add_adj_idcs else 0)
            x = G.placeholder([None, L * L * cin], "x")
            return N.growing_gen(x, cfg, use_batch_norm=cfg.batch_norm, currentUpres=cu, output=True, firstGen=True,
                                 filterSize=spec.filterSize, startFms=spec.startFms, maxFms=spec.maxFms,
                                 add_adj_idcs=spec.add_adj_idcs, first_nn_arch=spec.first_nn_arch,
                                 use_res_net=spec.use_res_net)
        if spec.add_adj_idcs:
            raise NotImplementedError("add_adj_idcs2/3 raise in the reference as well (App. D.7)")
        x = G.placeholder([None, L * L * C], "x")
        y = G.placeholder([None, S * S], "y")
        x_in = N.sampler_2_input(x, y, cfg)
        return N.growing_gen(x_in, cfg, use_batch_norm=cfg.batch_norm, currentUpres=cu, output=True, firstGen=False,
                             filterSize=spec.filterSize, startFms=spec.startFms, maxFms=spec.maxFms,
                             add_adj_idcs=False, first_nn_arch=False, use_res_net=spec.use_res_net)


def make_weights_out(L, seed, upRes=8, specs=None, nets=(1, 2), **cfg_kw):
    specs = specs or SHIPPED_8X
    cfg = N.config_out(L, upRes=upRes, **cfg_kw)
    out = {}
    for idx in nets:
        G.reset_default_graph()
        build_out_graph(idx, specs[idx], cfg)
        out[idx] = W.init_graph_variables(G.get_default_graph(), seed)
    return out


def auto_batches(S):
    """Slices per launch of generators 1, 2, 3 for S x S slices: ~8M / ~4M output pixels per launch (32 / 16 / 16 slices
    of 512^2, 2 / 1 / 1 of 2048^2); `_pick_batch` later reduces them to divisors of the rank's slice count."""
    px = int(S) * int(S)
    b1 = max(2, min(32, (1 << 23) // px))
    b23 = max(1, min(16, (1 << 22) // px))
    return (b1, b23, b23)


class MultiPassOut:
    """generate3DUniForNewNetwork (GAN/multipassGAN-out.py:390-618) for one frame, on the device."""

    def __init__(self, L, weights, upRes=8, specs=None, precision="fp16", transposeAxis=0, batches=None,
                 device=0, threshold=THRESHOLD, rank=0, world=1, group=None, range_check=False, **cfg_kw):
        """With world > 1 (generators 1+2, transposeAxis 0: the shipped 8x two-pass recipe) the volume is sharded by
        slice: z-slabs in pass 1, x-slabs in pass 2, one all-to-all per axis change; `__call__` then returns this
        rank's canonical z-slab [S/G, S, S]."""
        self.L, self.u, self.S = int(L), int(upRes), int(L) * int(upRes)
        self.rank, self.world, self.group = int(rank), int(world), group
        self.h = capi.default_handle(device)
        self.device = torch.device("cuda", device)
        self.specs = specs or SHIPPED_8X
        if batches is None:
            # The reference feeds 8 slices per sess.run to generator 1 and 2 to generators 2/3 (GAN/multipassGAN-out.py:
            # 439-447,501-509) because of its GPU memory; slices are independent at inference (no batch statistics), so the
            # batch only sets how many pixels one launch covers. ~8M output pixels per launch keep all 148 SMs busy for
            # many tiles at every stage (32 slices of 512^2, 2 of 2048^2) and bound the activations per layer.
            batches = auto_batches(self.S)
        self.cfg = N.config_out(self.L, upRes=self.u, **cfg_kw)
        self.ta = int(transposeAxis)
        self.threshold = float(threshold)
        self.nets = sorted(weights.keys())
        L, S, u = self.L, self.S, self.u
        f32 = dict(dtype=torch.float32, device=self.device)
        self.passes = {}
        self.flops = 0.0
        self.s0, self.s1 = par.slab_range(self.rank, self.world, S)
        self.S_loc = self.s1 - self.s0
        for idx in self.nets:
            spec = self.specs[idx]
            if self.ta not in _PASS_GEOM[idx]:
                raise NotImplementedError("transposeAxis %d is a dead branch for generator %d (App. D.7)" % (self.ta, idx))
            axis_of, chans = _PASS_GEOM[idx][self.ta]
            B = _pick_batch(self.S_loc, batches[idx - 1])
            pn = _PassNet(self.h, lambda idx=idx, spec=spec: build_out_graph(idx, spec, self.cfg), weights[idx], B,
                          precision, range_check)
            adj = bool(spec.add_adj_idcs) and idx == 1
            cin = 4 + (2 if adj else 0)
            desc = capi.make_assemble_desc((L, L, L), 4, axis_of, (u, 1, 1), chans, None, add_adj=adj,
                                           out_dtype=capi.F32, out_cstride=cin)
            self.passes[idx] = dict(net=pn, desc=desc, batch=B, inbuf=torch.empty((B, L, L, cin), **f32))
            self.flops += pn.net.flops / B * self.S_loc  # this rank's share
        self.vol_rows = torch.empty((self.S_loc, S, S), **f32)
        self.peer = par.make_peer_slab((self.S_loc, S, S), self.device, S, self.world, group)
        self.peer_c = par.make_peer_slab((self.S_loc, S, S), self.device, S, self.world, group) if self.peer else None
        if self.peer is not None and self.peer_c is None:
            self.peer = None
        self.vol_dim = self.peer.tensor if self.peer else torch.empty((self.S_loc, S, S), **f32)
        self.vol_out = self.peer_c.tensor if self.peer else (torch.empty((self.S_loc, S, S), **f32) if self.world > 1 else None)
        self.scr_a = self.scr_b = None
        if self.world > 1 and (self.peer is None or 3 in self.nets):
            # pack / receive staging of the NCCL all-to-all form of an axis change (no peer mapping, or the row-preserving
            # final permutation of the three-generator recipe, which the transposing peer-store kernel does not cover)
            self.scr_a = torch.empty((self.S_loc, S, S), **f32)
            self.scr_b = torch.empty((self.S_loc, S, S), **f32)
        self.launches_per_frame = sum((self.S_loc // p["batch"]) * (p["net"].net.launches + 1) for p in self.passes.values()) \
            + (len(self.nets) + 3 if self.world == 1 else 4)

    def _permute3(self, src, dst, dims, perm, thr):
        capi.transpose3d(self.h, src, dst, dims, perm, thr, torch.cuda.current_stream(self.device).cuda_stream)

    _POST = {1: (2, 1, 0), 2: (1, 2, 0), 3: (0, 1, 2)}  # GAN/multipassGAN-out.py:459, :521, :583

    def _axis_change(self, rows, perm, threshold, k):
        """k-th axis change of the frame on a slice-sharded volume: returns the buffer holding this rank's axis-0 slab of
        rows_full.transpose(perm). Targets alternate between two slabs, so nobody stores into memory a peer may still be
        reading in the current pass (parallel.PeerSlab); a permutation that keeps the slab axis is local."""
        S = self.S
        if perm == (0, 1, 2):
            if threshold > 0:
                capi.threshold(self.h, rows, self.S_loc * S * S, threshold, torch.cuda.current_stream(self.device).cuda_stream)
            return rows
        peer = (self.peer, self.peer_c)[k % 2] if self.peer else None
        out = peer.tensor if peer else (self.vol_dim, self.vol_out)[k % 2]
        if perm[0] == 0:
            self._permute3(rows, out, (self.S_loc, S, S), perm, threshold)
        elif peer is not None and perm[2] != 2:
            peer.exchange(capi, self.h, rows, S, perm[0], perm, threshold, wait_readers=False)
        else:
            par.reslab_any(rows, S, self.world, self.group, self._permute3, self.scr_a, self.scr_b, out, perm, threshold)
        return out

    def _call_sharded(self, x, output_free=None):
        """Any generator chain / transposeAxis on this rank's slices (App. C). After generator i the rows
        [slice_loc, row, col] are re-sliced into rows_full.transpose(post[i]) along ITS axis 0 (the `y` feed of generator
        i+1, GAN/multipassGAN-out.py:459,521); after the last one post[i] and the closing transposes (:587-590) are
        composed into ONE axis change with the threshold (:612-615) fused. Returns this rank's canonical z-slab."""
        S = self.S
        cur = torch.cuda.current_stream(self.device)
        st = cur.cuda_stream
        vol = _dev_f32(x, self.device)
        rows = self.vol_rows
        dim = None
        closing = []
        if 2 in self.nets:
            closing.append((2, 0, 1))
        if 1 in self.nets:
            closing.append((2, 1, 0))
        if output_free is not None and len(self.nets) != 2:
            cur.wait_event(output_free)  # the finished frame may live in the slab the first pass already pushes into
            output_free = None
        for k, idx in enumerate(self.nets):
            p = self.passes[idx]
            B = p["batch"]
            last = idx == self.nets[-1]
            perm = par.compose_perms(self._POST[idx], *closing) if last else self._POST[idx]
            # with peers every finished batch of the FIRST pass is pushed to its owners at once (4-aligned batches when
            # the part's rows become the contiguous axis: 128-bit stores)
            stream_parts = (self.peer is not None and k == 0 and not last and perm[2] != 2 and perm[0] != 0
                            and (perm[2] != 0 or B % 4 == 0))
            for s in range(self.s0, self.s1, B):
                capi.slice_assemble(self.h, p["desc"], vol, None, s, B, p["inbuf"], st)
                feeds = {"x": p["inbuf"]}
                if idx > 1:
                    feeds["y"] = dim[s - self.s0]
                p["net"].net.run(feeds, out=rows[s - self.s0], stream=st)
                if stream_parts:
                    self.peer.push_part(capi, self.h, rows[s - self.s0], S, s, B, perm[0], perm, 0.0)
            if k == 0 and output_free is not None:
                cur.wait_event(output_free)
            if stream_parts:
                dim = self.peer.landed(channel=0)
            else:
                dim = self._axis_change(rows, perm, self.threshold if last else 0.0, k)
        return dim

    def __call__(self, x, output_free=None):
        """x: [L,L,L,4] float32, velocities already scaled by velScale (GAN/multipassGAN-out.py:138).
        output_free: see MultiPass4x.__call__ (sharded runs; the single-GPU path ping-pongs its two volumes)."""
        if self.world > 1:
            return self._call_sharded(x, output_free)
        if output_free is not None:
            torch.cuda.current_stream(self.device).wait_event(output_free)
        S = self.S
        st = torch.cuda.current_stream(self.device).cuda_stream
        vol = _dev_f32(x, self.device)
        rows, dim = self.vol_rows, self.vol_dim
        post = {1: (2, 1, 0), 2: (1, 2, 0), 3: (0, 1, 2)}  # :459, :521, :583
        for idx in self.nets:
            p = self.passes[idx]
            B = p["batch"]
            for s0 in range(0, S, B):
                capi.slice_assemble(self.h, p["desc"], vol, None, s0, B, p["inbuf"], st)
                feeds = {"x": p["inbuf"]}
                if idx > 1:
                    feeds["y"] = dim[s0]
                p["net"].net.run(feeds, out=rows[s0], stream=st)
            # np.array(rows).reshape(S,S,S).transpose(post): the old `dim` (this pass's `y`) is dead now
            capi.transpose3d(self.h, rows, dim, (S, S, S), post[idx], 0.0, st)
        cur = dim
        other = rows
        # :587-590 -- undo the pass transposes
        if 2 in self.nets:
            capi.transpose3d(self.h, cur, other, (S, S, S), (2, 0, 1), 0.0, st)
            cur, other = other, cur
        if 1 in self.nets:
            capi.transpose3d(self.h, cur, other, (S, S, S), (2, 1, 0), 0.0, st)
            cur, other = other, cur
        if self.threshold > 0:
            capi.threshold(self.h, cur, S * S * S, self.threshold, st)  # :612-615
        self.vol_rows, self.vol_dim = other, cur
        return cur
