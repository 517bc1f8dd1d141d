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
