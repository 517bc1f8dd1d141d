"""ORACLE (test infrastructure only): numpy restatement of the tile cut / stitch of
tools_wscale/tilecreator_t.py (createTiles :403-434, cutTile :436-450, concatTiles :886-918) for 2-D slices.
PINNED bit-exactly against the reference's own methods executed by tests/golden/make_golden.py (tiles.npz)."""
import numpy as np


def create_tiles(data, tile_shape, strides=-1, padding=0):
    """data: [z=1, H, W, C]; tile_shape [1, th, tw]. Returns [tiles, 1(+0 pad), th+2p, tw+2p, C]."""
    ds = data.shape
    pad = [padding, padding, padding, 0]
    if np.isscalar(strides):
        strides = list(tile_shape) if strides <= 0 else [strides, strides, strides]
    strides = list(strides)
    if ds[0] <= 1:
        pad[0] = 0
        strides[0] = 1
    n = [(ds[i] - tile_shape[i]) // strides[i] + 1 for i in range(3)]
    tiles = []
    for tz in range(n[0]):
        for ty in range(n[1]):
            for tx in range(n[2]):
                f = [tz * strides[0], ty * strides[1], tx * strides[2]]
                t = data[f[0]:f[0] + tile_shape[0], f[1]:f[1] + tile_shape[1], f[2]:f[2] + tile_shape[2], :]
                if padding > 0:  # the reference passes the flat list [p,p,p,0], which numpy rejects; evident intent:
                    t = np.pad(t, [(p, p) for p in pad], "edge")
                tiles.append(t)
    return np.array(tiles)


def concat_tiles(tiles, frame_shape, tile_border=(0, 0, 0, 0)):
    """tiles [batch, z, y, x, c]; frame_shape in tiles [z, y, x]; crop tile_border [z,y,x,c] from both sides."""
    b = np.asarray(tile_border)
    if (b > 0).any():
        shp = np.asarray(tiles.shape[1:]) - 2 * b
        tiles = [t[b[0]:b[0] + shp[0], b[1]:b[1] + shp[1], b[2]:b[2] + shp[2], :] for t in tiles]
    frame = []
    for z in range(frame_shape[0]):
        rows = []
        for y in range(frame_shape[1]):
            off = z * frame_shape[1] * frame_shape[2] + y * frame_shape[2]
            rows.append(np.concatenate(tiles[off:off + frame_shape[2]], axis=2))
        frame.append(np.concatenate(rows, axis=1))
    return np.concatenate(frame, axis=0)


def cut_overlap(frames, tile, border):
    """frames [n, H, W, C] with H = ty*(tile-2*border)+2*border (same for W): overlapped cut with stride tile-2*border,
    no padding (what mpg_tiles_cut does for the tiled apply). Returns [n*ty*tx, tile, tile, C]."""
    core = tile - 2 * border
    n, H, W, _ = frames.shape
    ty, tx = (H - 2 * border) // core, (W - 2 * border) // core
    assert ty * core + 2 * border == H and tx * core + 2 * border == W
    return np.stack([frames[i, y * core:y * core + tile, x * core:x * core + tile] for i in range(n) for y in range(ty)
                     for x in range(tx)])


def stitch_overlap(tiles, n, ty, tx, border):
    """Restatement of mpg_tiles_stitch_overlap (NOT a reference function: concatTiles drops the frame-edge band, this
    keeps it): tiles [n*ty*tx, th, tw, C] -> [n, ty*(th-2b)+2b, tx*(tw-2b)+2b, C]. Every tile contributes its centre,
    tiles on a frame edge also their outer border."""
    th, tw, c = tiles.shape[1:]
    ch, cw = th - 2 * border, tw - 2 * border
    out = np.empty((n, ty * ch + 2 * border, tx * cw + 2 * border, c), tiles.dtype)
    for i in range(n):
        for y in range(ty):
            y0 = 0 if y == 0 else border
            y1 = th if y == ty - 1 else th - border
            for x in range(tx):
                x0 = 0 if x == 0 else border
                x1 = tw if x == tx - 1 else tw - border
                t = tiles[(i * ty + y) * tx + x]
                out[i, y * ch + y0:y * ch + y1, x * cw + x0:x * cw + x1] = t[y0:y1, x0:x1]
    return out
