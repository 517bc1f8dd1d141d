"""Training batch source: `TileCreator.addData` + `selectRandomTiles(augment=False)` of
tools_wscale/tilecreator_t.py (:321-392, :457-489, :548-646, :920-926) for 2-D data (dim=2, dim_t=1) -- the default of
GAN/multipassGAN-4x.py (`dataAugmentation 0`, :95; `getinput` :1017-1047 reshapes the tiles to rows).

The reference cuts every tile with numpy on the host and feeds the batch through `feed_dict` each step. Here the frames
live where the trainer lives (a CUDA device, or the CPU in tests): the host only replays the reference's DECISIONS --
the same `random.randrange` call sequence (frame, then up to 19 offset tries of three calls each, including the
degenerate `randrange(0, 1)` of the z axis, which consumes generator state) and the same float64 density test on the
low-res tile -- and one gather per batch builds `[n, T*T*C]` / `[n, (T*u)^2]` rows directly in device memory.
Pinned against the reference's own methods: tests/golden/tilesampler.npz (tests/golden/make_golden.py sampler).

Data augmentation (`selectRandomTiles(augment=True)` -> `generateTile`, :491-546) as the shipped 4x command line uses it
(GAN/example_run_output.py:6: `dataAugmentation 1 rot 1`; defaults minScale 0.85, maxScale 1.15, flip 1): random scaling
(scipy.ndimage.zoom order 1 = align-corners bilinear), a second random cut, a random 90-degree rotation and a random flip
with the velocity channels fixed up (:755-879). Same split: the host replays the decisions (Python `random.randrange` for
frame / offsets, numpy's legacy `RandomState` for the scale factor, rotation and flip -- the reference calls the global
`np.random.uniform` / `np.random.choice`), the device does the data movement and interpolation. Pinned by
tests/golden/tileaugment.npz (the reference's own generateTile, tests/golden/make_golden.py augment).
The free-angle rotation (`rot 2`: scipy affine_transform) is not ported.
"""
import random

import numpy as np
import torch


class TileSamplerError(Exception):
    """tilecreator_t.py:1063-1067 TilecreatorError."""


class TileSampler:
    def __init__(self, tileSizeLow, upres, densityMinimum=0.02, partTrain=0.9, partTest=0.1, partVal=0, device=None,
                 rng=None, dim_t=1):
        """rng: an object with `randrange(a, b)` (default: the `random` module, like the reference's
        `from random import randrange`); pass `random.Random(seed)` for reproducible batches.
        dim_t: frames per datum (TileCreator(dim_t=3) of the 8x trainer: the frames of a sequence are channel groups,
        low [N,1,L,L,C*dim_t], high [N,1,S,S,Ch*dim_t]); tiles then carry all dim_t frames (`tile_t = dim_t`, as
        selectRandomTempoTiles asks for, :1391) and the augmentation's velocity fix-ups act on every frame's group."""
        self.T, self.u = int(tileSizeLow), int(upres)
        self.dim_t = int(dim_t)
        self.density_minimum = float(densityMinimum)
        total = partTrain + partTest + partVal
        self.part_train, self.part_test = partTrain / total, partTest / total  # :222-225
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.rng = rng if rng is not None else random
        self.low = None     # [N, L, L, C]  on self.device
        self.high = None    # [N, S, S, Ch]
        self._dens = None   # host copy of the low-res density channel of every frame, float32 [N, dim_t, L, L] (density test)
        self.set_borders = [0, 0, 0]
        self.use_data_aug = False

    def init_data_augmentation(self, rot=2, minScale=0.85, maxScale=1.15, flip=True, np_rng=None):
        """TileCreator.initDataAugmentation (:227-317). rot: 1 = 90-degree rotations, 2 = free rotation (not ported),
        else none; minScale == maxScale == 1 disables scaling. np_rng: a numpy RandomState (the reference draws from the
        global one; `np.random.RandomState(seed)` reproduces `np.random.seed(seed)`)."""
        if rot == 2:
            raise NotImplementedError("free-angle rotation (rot 2, scipy affine_transform) is not ported; the shipped 4x "
                                      "recipe trains with rot 1 (GAN/example_run_output.py:6)")
        self.use_data_aug = True
        self.do_rot90 = rot == 1
        self.scale_factor = (float(minScale), float(maxScale))
        self.do_scaling = not (minScale == 1 and maxScale == 1)
        self.do_flip = bool(flip)
        self.np_rng = np_rng if np_rng is not None else np.random
        return self

    # ------------------------------------------------------------------ data
    def add_data(self, low, high):
        """low: [N, 1, L, L, C] (or [1, L, L, C]), high: [N, 1, S, S, Ch]; appended like addData (:321-377)."""
        low, high = np.asarray(low, np.float32), np.asarray(high, np.float32)
        if low.ndim != high.ndim:
            raise TileSamplerError("Data shape mismatch. Dimensions: %d vs %d" % (low.ndim, high.ndim))
        if low.ndim == 4:
            low, high = low[None], high[None]
        if low.ndim != 5:
            raise TileSamplerError("Input must be single 3D data or sequence of 3D data.")
        if low.shape[0] != high.shape[0]:
            raise TileSamplerError("unequal amount of low (%d) and high (%d) data." % (low.shape[0], high.shape[0]))
        if low.shape[1] != 1 or high.shape[1] != 1:
            raise TileSamplerError("only 2-D data (z extent 1) is supported")
        L, S = low.shape[2], high.shape[2]
        if low.shape[3] != L or high.shape[3] != S or S != L * self.u:
            raise TileSamplerError("Frame shape mismatch: low %s high %s upres %d" % (low.shape, high.shape, self.u))
        if L < self.T:
            raise TileSamplerError("Can't cut tile %d from frame %d." % (self.T, L))
        lo = torch.from_numpy(np.ascontiguousarray(low[:, 0])).to(self.device)
        hi = torch.from_numpy(np.ascontiguousarray(high[:, 0])).to(self.device)
        if low.shape[-1] % self.dim_t or high.shape[-1] % self.dim_t:
            raise TileSamplerError("channel counts %d / %d are not multiples of dim_t %d" % (low.shape[-1], high.shape[-1], self.dim_t))
        C = low.shape[-1] // self.dim_t
        # density channel of every frame of the datum, [N, dim_t, L, L] (the density test sees channel 0 of the CUT tile, i.e.
        # of the first frame the tile carries)
        dens = np.ascontiguousarray(np.stack([low[:, 0, :, :, k * C] for k in range(self.dim_t)], axis=1))
        if self.low is None:
            self.low, self.high, self._dens = lo, hi, dens
        else:
            if tuple(self.low.shape[1:]) != tuple(lo.shape[1:]) or tuple(self.high.shape[1:]) != tuple(hi.shape[1:]):
                raise TileSamplerError("Frame shape mismatch with the data already added")
            self.low, self.high = torch.cat([self.low, lo]), torch.cat([self.high, hi])
            self._dens = np.concatenate([self._dens, dens])
        n = self.low.shape[0]
        end_train = int(n * self.part_train)  # splitSets :379-388
        self.set_borders = [end_train, end_train + int(n * self.part_test), n]

    # ------------------------------------------------------------------ the reference's decisions
    def _tile_t(self, tile_t):
        """Frames per tile: None = the whole sequence (tile_t = dim_t, selectRandomTempoTiles :1391); 1 = the reference's
        default for getinput, which on dim_t > 1 data also draws WHICH frame (getRandomDatum :548-560)."""
        t = self.dim_t if tile_t is None else int(tile_t)
        if t > self.dim_t:
            raise TileSamplerError("not enough coherent frames. Requested %d, given %d" % (t, self.dim_t))
        return t

    def _rand_frame(self, tile_t):
        """getRandomDatum :555-560: `randrange(0, dim_t - tile_t)` when tile_t < dim_t (so the last frame is never the first of
        a tile), no draw otherwise."""
        return self.rng.randrange(0, self.dim_t - tile_t) if tile_t < self.dim_t else 0

    def select_offsets(self, selection_size, is_training=True, tile_t=None):
        """[(frame, oy, ox, first frame of the sequence)] * selection_size, consuming the generator exactly like
        selectRandomTiles (:457-489)."""
        if is_training:
            if self.set_borders[0] < 1:
                raise TileSamplerError("no training data.")
        elif self.set_borders[1] - self.set_borders[0] < 1:
            raise TileSamplerError("no test data.")
        L, T = self._dens.shape[2], self.T
        tile_t = self._tile_t(tile_t)
        end = L - T + 1
        need = self.density_minimum * 1 * T * T  # hasMinDensity :920-921
        rr = self.rng.randrange
        picks = []
        for _ in range(int(selection_size)):
            f = rr(0, self.set_borders[0]) if is_training else rr(self.set_borders[0], self.set_borders[1])  # :548-553
            fr = self._rand_frame(tile_t)
            i, ok = 1, False
            oy = ox = 0
            while (not ok) and i < 20:  # getRandomTile :622-640
                rr(0, 1)  # the z offset of 2-D data: always 0, but the call advances the generator
                oy = rr(0, end)
                ox = rr(0, end)
                ok = float(self._dens[f, fr, oy:oy + T, ox:ox + T].sum(dtype=np.float64)) >= need  # getTileDensity :923-926
                i += 1
            picks.append((f, oy, ox, fr))
        return picks

    # ------------------------------------------------------------------ gather
    def gather(self, picks, tile_t=None):
        """-> (low [n, 1, T, T, C*tile_t], high [n, 1, T*u, T*u, Ch*tile_t]) tensors on self.device (one advanced-indexing
        gather each, then the channel groups of the picked frames)."""
        T, u = self.T, self.u
        dev = self.device
        tile_t = self._tile_t(tile_t)
        f = torch.tensor([p[0] for p in picks], device=dev, dtype=torch.long)
        oy = torch.tensor([p[1] for p in picks], device=dev, dtype=torch.long)
        ox = torch.tensor([p[2] for p in picks], device=dev, dtype=torch.long)
        ar = torch.arange(T, device=dev)
        low = self.low[f[:, None, None], (oy[:, None] + ar)[:, :, None], (ox[:, None] + ar)[:, None, :]]
        aru = torch.arange(T * u, device=dev)
        high = self.high[f[:, None, None], (oy[:, None] * u + aru)[:, :, None], (ox[:, None] * u + aru)[:, None, :]]
        if tile_t < self.dim_t:  # getDatum :562-573: the channel groups of frames [fr, fr + tile_t)
            C, Ch = low.shape[-1] // self.dim_t, high.shape[-1] // self.dim_t
            fr = torch.tensor([p[3] for p in picks], device=dev, dtype=torch.long)
            low = torch.gather(low, 3, (fr[:, None] * C + torch.arange(C * tile_t, device=dev))[:, None, None, :].expand(-1, T, T, -1))
            high = torch.gather(high, 3, (fr[:, None] * Ch + torch.arange(Ch * tile_t, device=dev))[:, None, None, :].expand(-1, T * u, T * u, -1))
        return low.unsqueeze(1), high.unsqueeze(1)

    def select_random_tiles(self, selection_size, is_training=True, augment=False, tile_t=None):
        """selectRandomTiles(selectionSize, isTraining, augment, tile_t): (batch_low, batch_high). tile_t: see _tile_t (None =
        every frame of the datum; the reference's own default is 1)."""
        if augment and self.use_data_aug:
            return self.generate_tiles(selection_size, is_training, tile_t)
        return self.gather(self.select_offsets(selection_size, is_training, tile_t), tile_t)

    # ------------------------------------------------------------------ augmentation (generateTile :491-546)
    def _random_offset(self, f, frame_h, tile, dens_of):
        """getRandomTile (:574-640) on a square 2-D frame: up to 19 tries of (z, y, x) offsets, density test on the cut."""
        end = frame_h - tile + 1
        if end < 1:
            raise TileSamplerError("Can't cut tile %d from frame %d." % (tile, frame_h))
        need = self.density_minimum * 1 * tile * tile
        rr = self.rng.randrange
        i, ok = 1, False
        oy = ox = 0
        while (not ok) and i < 20:
            rr(0, 1)
            oy = rr(0, end)
            ox = rr(0, end)
            ok = dens_of(oy, ox, tile) >= need
            i += 1
        return oy, ox

    def generate_tiles(self, selection_size, is_training=True, tile_t=None):
        """`selection_size` augmented (low, high) tile pairs: [n, 1, T, T, C] / [n, 1, T*u, T*u, Ch] on self.device."""
        import torch.nn.functional as F
        T, u = self.T, self.u
        L = self._dens.shape[2]
        rr = self.rng.randrange
        tile_t = self._tile_t(tile_t)
        lows, highs = [], []
        for _ in range(int(selection_size)):
            f = rr(0, self.set_borders[0]) if is_training else rr(self.set_borders[0], self.set_borders[1])
            fr = self._rand_frame(tile_t)
            low, high = self.low[f], self.high[f]  # [L,L,C], [S,S,Ch] views on the device
            if tile_t < self.dim_t:
                C, Ch = low.shape[-1] // self.dim_t, high.shape[-1] // self.dim_t
                low, high = low[..., fr * C:(fr + tile_t) * C], high[..., fr * Ch:(fr + tile_t) * Ch]
            dens = self._dens[f, fr]
            if self.do_scaling:
                sf = float(self.np_rng.uniform(self.scale_factor[0], self.scale_factor[1]))
                tb = int(np.ceil(T * (1.0 / sf)))  # tile cut "for faster transformation" (:503-512)
                oy, ox = self._random_offset(f, L, tb, lambda y, x, t: float(dens[y:y + t, x:x + t].sum(dtype=np.float64)))
                low = low[oy:oy + tb, ox:ox + tb]
                high = high[oy * u:(oy + tb) * u, ox * u:(ox + tb) * u]
                # scale (:818-851): zoom factors round(shape_low * f) / shape_low for BOTH tensors, order-1 zoom
                ts = int(np.round(tb * sf))
                zf = ts / tb
                hs = int(round(tb * u * zf))
                low = F.interpolate(low.permute(2, 0, 1)[None], size=(ts, ts), mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
                high = F.interpolate(high.permute(2, 0, 1)[None], size=(hs, hs), mode="bilinear", align_corners=True)[0].permute(1, 2, 0)
                low = self._vel_map(low, lambda v: [c * np.float32(sf) for c in v])  # scaleVelocities :853-862
                dens_t = low[..., 0].double().cpu().numpy()  # the density test of the second cut sees the scaled tile
                oy, ox = self._random_offset(f, ts, T, lambda y, x, t: float(dens_t[y:y + t, x:x + t].sum(dtype=np.float64)))
            else:
                oy, ox = self._random_offset(f, L, T, lambda y, x, t: float(dens[y:y + t, x:x + t].sum(dtype=np.float64)))
            low = low[oy:oy + T, ox:ox + T]
            high = high[oy * u:(oy + T) * u, ox * u:(ox + T) * u]
            if self.do_rot90:  # cube_rot[2] = [[], [z], [z, z], [nz]], z = (2, 1), nz = (1, 2) (:292-300)
                k = int(self.np_rng.randint(0, 4))
                for axes in ([], [(2, 1)], [(2, 1), (2, 1)], [(1, 2)])[k]:
                    # np.rot90(data[1,y,x,c], axes=(a0,a1)) on the tile without its unit z axis: axes (1,2) -> (0,1)
                    dims = (axes[0] - 1, axes[1] - 1)
                    low, high = torch.rot90(low, 1, dims), torch.rot90(high, 1, dims)
                    # Reference quirk (reproduced, not fixed): rotate90Velocities (:755-763) swaps entries of a LIST of
                    # channel views and returns a new array, and special_aug (:648-663) only writes an op's result back
                    # for tile_t > 1 -- so with single frames the velocity vectors are NOT rotated with the tile
                    # (flipVelocities / scaleVelocities modify their views in place and do take effect).
                    if tile_t > 1:  # multi-frame tiles: the result IS written back, every frame's vectors rotate
                        a0, a1 = 2 - axes[0], 2 - axes[1]  # axes z,y,x = 0,1,2 -> velocity components x,y,z = 0,1,2

                        def rot(v, a0=a0, a1=a1):
                            v = list(v)
                            v[a0], v[a1] = -v[a1], v[a0]
                            return v
                        low = self._vel_map(low, rot)
            if self.do_flip:
                axis = int(self.np_rng.randint(0, 4))
                if axis < 3:
                    if axis > 0:
                        low, high = torch.flip(low, (axis - 1,)), torch.flip(high, (axis - 1,))
                    def flip_v(v, k=2 - axis):  # flipVelocities (:797-813): axis 2 -> vx, 1 -> vy, 0 -> vz
                        v = list(v)
                        v[k] = -v[k]
                        return v
                    low = self._vel_map(low, flip_v)
            if tuple(low.shape[:2]) != (T, T) or tuple(high.shape[:2]) != (T * u, T * u):
                raise TileSamplerError("Wrong tile shape after data augmentation. is: %s,%s." % (tuple(low.shape), tuple(high.shape)))
            lows.append(low)
            highs.append(high)
        return torch.stack(lows).unsqueeze(1).contiguous(), torch.stack(highs).unsqueeze(1).contiguous()

    def _vel_map(self, low, fn):
        """Apply fn([vx, vy, vz]) -> [vx', vy', vz'] to the velocity channels (1, 2, 3) of every frame's channel group
        (special_aug :648-663 reshapes multi-frame data to [-1, tile_t, channels] before it calls the op)."""
        ch = list(low.unbind(-1))
        C = self.low.shape[-1] // self.dim_t
        for k in range(len(ch) // C):
            b = k * C
            ch[b + 1], ch[b + 2], ch[b + 3] = fn([ch[b + 1], ch[b + 2], ch[b + 3]])
        return torch.stack(ch, dim=-1)

    def batch_rows(self, batch_size, is_training=True, augment=False, tile_t=None):
        """getinput (GAN/multipassGAN-4x.py:1017-1047) with useVelocities and no vorticity / velocity modification:
        (batch_xs [n, T*T*C], batch_ys [n, (T*u)^2 * Ch]) in device memory, ready for Trainer4x.iteration."""
        low, high = self.select_random_tiles(batch_size, is_training, augment, tile_t)
        n = low.shape[0]
        return low.reshape(n, -1), high.reshape(n, -1)
