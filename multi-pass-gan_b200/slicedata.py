"""3-D volume -> 2-D training slices (second half): the `conv_slices` path of FluidDataLoader.loadFiles
(tools_wscale/fluiddataloader.py:414-431 axis conversion + velocity-channel swap, :511-524 axis zoom / removeSlices /
selectRandomSamples, helpers :295-347) that feeds TileCreator.addData in the multi-pass training scripts
(GAN/multipassGAN-4x.py:238-262: conv_axis / axis_scaling per upsampling mode).

The reference does this with numpy / scipy on the host for every loaded file; here every step is a torch op on the
tensor's own device (the frames then stay resident for tilesampler.TileSampler), computed once per volume.
Reference quirks that change results are reproduced:
  * selectRandomSamples (:337-347) calls `np.random.shuffle` on a temporary and indexes with its None result, so NO
    shuffle happens: the first int(n * select_random) surviving slices are kept, in order.
  * the channel swap loops over three 4-channel groups (tempo data); groups beyond the tensor's channels are empty.
Pinned against the reference's own statements / methods: tests/golden/slicedata.npz (tests/golden/make_golden.py slices).
"""
import torch


def convert_slices(fx, conv_axis):
    """fx [Z,Y,X,C] -> slices along `conv_axis` first (fluiddataloader.py:414-431): 1: transpose(1,0,2,3) and swap
    channels 2<->3 of every (d,vx,vy,vz) group; 2: transpose(2,1,0,3) and swap channels 1<->3; 0: unchanged."""
    if conv_axis not in (0, 1, 2):
        raise ValueError("conv_axis must be 0, 1 or 2")
    if conv_axis == 0:
        return fx
    fx = fx.permute(1, 0, 2, 3) if conv_axis == 1 else fx.permute(2, 1, 0, 3)
    c = fx.shape[3]
    if c > 3:
        idx = list(range(c))
        a, b = (2, 3) if conv_axis == 1 else (1, 3)
        for i in range(3):
            if i * 4 + max(a, b) < c:
                idx[i * 4 + a], idx[i * 4 + b] = idx[i * 4 + b], idx[i * 4 + a]
        fx = fx[..., idx]
    return fx.contiguous()


def zoom_linear(a, zoom):
    """scipy.ndimage.zoom(a, zoom, order=1) for the per-axis factors of `axis_scaling` (:513-514): output extent
    round(n * z), coordinates o * (n - 1) / (n_out - 1) (align-corners), linear interpolation along each scaled axis."""
    for ax, z in enumerate(zoom):
        if float(z) == 1.0:
            continue
        n = a.shape[ax]
        n_out = int(round(n * float(z)))
        if n_out <= 1:
            a = a.narrow(ax, 0, 1)
            continue
        pos = torch.arange(n_out, device=a.device, dtype=torch.float64) * ((n - 1) / (n_out - 1))
        lo = pos.floor().clamp(max=n - 1).long()
        hi = (lo + 1).clamp(max=n - 1)
        w = (pos - lo.double()).to(a.dtype)
        shape = [1] * a.dim()
        shape[ax] = n_out
        w = w.view(shape)
        a = a.index_select(ax, lo) * (1 - w) + a.index_select(ax, hi) * w
    return a


def add_adj_slices(fx):
    """addAdjSlices (:314-335): per 4-channel group append the density of slice i-1 and i+1 (zeros at the ends)."""
    n, h, w, _ = fx.shape
    g = fx.reshape(n, h, w, 3, -1)
    prev = torch.zeros_like(g[..., 0:1])
    nxt = torch.zeros_like(g[..., 0:1])
    prev[1:] = g[:-1, ..., 0:1]
    nxt[:-1] = g[1:, ..., 0:1]
    return torch.cat([g, prev, nxt], dim=-1).reshape(n, h, w, -1)


def remove_slices(fx, fy=None, density_threshold=0.002):
    """removeSlices (:295-312): keep the slices whose mean density (channel 0) reaches the threshold, in order."""
    keep = fx[..., 0:1].float().mean(dim=(1, 2, 3)) >= density_threshold
    idx = keep.nonzero().flatten()
    if fy is None:
        return fx.index_select(0, idx)
    return fx.index_select(0, idx), fy.index_select(0, idx)


def select_random_samples(fx, fy, select_random):
    """selectRandomSamples (:337-347) as it behaves: the first int(n * select_random) slices (see module docstring)."""
    k = int(fx.shape[0] * select_random)
    return fx[:k], fy[:k]


def slices_from_volumes(fx, fy, conv_axis=0, axis_scaling=(1, 1, 1, 1), axis_scaling_y=(1, 1, 1, 1), density_threshold=0.002,
                        select_random=0.1, add_adj_idcs=False):
    """One (low, high) volume pair -> the 2-D frames FluidDataLoader hands to TileCreator.addData (conv_slices=True,
    have_y): fx [Z,Y,X,C], fy [Zh,Yh,Xh,Ch] tensors (any device) -> (x [n,Y',X',C'], y [n,Yh',Xh',Ch])."""
    fx, fy = convert_slices(fx, conv_axis), convert_slices(fy, conv_axis)
    fx, fy = zoom_linear(fx, axis_scaling), zoom_linear(fy, axis_scaling_y)
    if add_adj_idcs:
        fx = add_adj_slices(fx)
    if fx.shape[0] != fy.shape[0]:
        raise ValueError("slice counts differ after axis scaling: %d vs %d" % (fx.shape[0], fy.shape[0]))
    fx, fy = remove_slices(fx, fy, density_threshold)
    return select_random_samples(fx, fy, select_random)


def frame_indices(filename_index_min, filename_index_max, data_fraction):
    """Which file indices FluidDataLoader loads from an index range (fluiddataloader.py:238-244, the "simple index range"
    branch): n = max(1, int((max - min) * data_fraction)) indices spread evenly over the range, int(min + t * (max - min) / n).
    Deterministic -- the loader's numpy seed does not enter.  Pinned by tests/golden/tempotiles.npz (frac_*)."""
    lo, hi = int(filename_index_min), int(filename_index_max)
    n = max(1, int((hi - lo) * data_fraction))
    tf = float(hi - lo) / n
    return [int(lo + t * tf) for t in range(n)]
