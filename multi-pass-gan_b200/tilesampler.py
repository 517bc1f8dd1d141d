"""Training batch source: `TileCreator.addData` + `selectRandomTiles(augment=False)` of
tools_wscale/tilecreator_t.py (:321-392, :457-489, :548-646, :920-926) for 2-D data (dim=2, dim_t=1) -- the default of
GAN/multipassGAN-4x.py (`dataAugmentation 0`, :95; `getinput` :1017-1047 reshapes the tiles to rows).

The reference cuts every tile with numpy on the host and feeds the batch through `feed_dict` each step. Here the frames
live where the trainer lives (a CUDA device, or the CPU in tests): the host only replays the reference's DECISIONS --
the same `random.randrange` call sequence (frame, then up to 19 offset tries of three calls each, including the
degenerate `randrange(0, 1)` of the z axis, which consumes generator state) and the same float64 density test on the
low-res tile -- and one gather per batch builds `[n, T*T*C]` / `[n, (T*u)^2]` rows directly in device memory.
Pinned against the reference's own methods: tests/golden/tilesampler.npz (tests/golden/make_golden.py sampler).
Augmentation (scipy rotations / scaling, `generateTile`) is not ported.
"""
import random

import numpy as np
import torch


class TileSamplerError(Exception):
    """tilecreator_t.py:1063-1067 TilecreatorError."""


class TileSampler:
    def __init__(self, tileSizeLow, upres, densityMinimum=0.02, partTrain=0.9, partTest=0.1, partVal=0, device=None,
                 rng=None):
        """rng: an object with `randrange(a, b)` (default: the `random` module, like the reference's
        `from random import randrange`); pass `random.Random(seed)` for reproducible batches."""
        self.T, self.u = int(tileSizeLow), int(upres)
        self.density_minimum = float(densityMinimum)
        total = partTrain + partTest + partVal
        self.part_train, self.part_test = partTrain / total, partTest / total  # :222-225
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.rng = rng if rng is not None else random
        self.low = None     # [N, L, L, C]  on self.device
        self.high = None    # [N, S, S, Ch]
        self._dens = None   # host copy of the low-res density channel, float32 [N, L, L] (density test)
        self.set_borders = [0, 0, 0]

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
        dens = np.ascontiguousarray(low[:, 0, :, :, 0])
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
    def select_offsets(self, selection_size, is_training=True):
        """[(frame, oy, ox)] * selection_size, consuming the generator exactly like selectRandomTiles (:457-489)."""
        if is_training:
            if self.set_borders[0] < 1:
                raise TileSamplerError("no training data.")
        elif self.set_borders[1] - self.set_borders[0] < 1:
            raise TileSamplerError("no test data.")
        L, T = self._dens.shape[1], self.T
        end = L - T + 1
        need = self.density_minimum * 1 * T * T  # hasMinDensity :920-921
        rr = self.rng.randrange
        picks = []
        for _ in range(int(selection_size)):
            f = rr(0, self.set_borders[0]) if is_training else rr(self.set_borders[0], self.set_borders[1])  # :548-553
            i, ok = 1, False
            oy = ox = 0
            while (not ok) and i < 20:  # getRandomTile :622-640
                rr(0, 1)  # the z offset of 2-D data: always 0, but the call advances the generator
                oy = rr(0, end)
                ox = rr(0, end)
                ok = float(self._dens[f, oy:oy + T, ox:ox + T].sum(dtype=np.float64)) >= need  # getTileDensity :923-926
                i += 1
            picks.append((f, oy, ox))
        return picks

    # ------------------------------------------------------------------ gather
    def gather(self, picks):
        """-> (low [n, 1, T, T, C], high [n, 1, T*u, T*u, Ch]) tensors on self.device (one advanced-indexing gather each)."""
        T, u = self.T, self.u
        dev = self.device
        f = torch.tensor([p[0] for p in picks], device=dev, dtype=torch.long)
        oy = torch.tensor([p[1] for p in picks], device=dev, dtype=torch.long)
        ox = torch.tensor([p[2] for p in picks], device=dev, dtype=torch.long)
        ar = torch.arange(T, device=dev)
        low = self.low[f[:, None, None], (oy[:, None] + ar)[:, :, None], (ox[:, None] + ar)[:, None, :]]
        aru = torch.arange(T * u, device=dev)
        high = self.high[f[:, None, None], (oy[:, None] * u + aru)[:, :, None], (ox[:, None] * u + aru)[:, None, :]]
        return low.unsqueeze(1), high.unsqueeze(1)

    def select_random_tiles(self, selection_size, is_training=True):
        """selectRandomTiles(selectionSize, isTraining, augment=False): (batch_low, batch_high)."""
        return self.gather(self.select_offsets(selection_size, is_training))

    def batch_rows(self, batch_size, is_training=True):
        """getinput (GAN/multipassGAN-4x.py:1017-1047) with useVelocities and no vorticity / velocity modification:
        (batch_xs [n, T*T*C], batch_ys [n, (T*u)^2 * Ch]) in device memory, ready for Trainer4x.iteration."""
        low, high = self.select_random_tiles(batch_size, is_training)
        n = low.shape[0]
        return low.reshape(n, -1), high.reshape(n, -1)
