"""mantaflow `.uni` grid files (what the reference reads and writes through tools_wscale/uniio.py:28-123).

Layout: gzip stream = 4-byte id (`MNT3`: dimX dimY dimZ gridType elementType bytesPerElement info[252] dimT
timestamp = struct 'iiiiii252siQ'; `MNT2`: same with info[256] and no dimT = 'iiiiii256sQ'), followed by the
float32 / int32 payload in C order [Z, Y, X, channels] (vec3 grids: elementType 2, 12 bytes per element).
"""
import gzip
import struct

import numpy as np

_V4 = "iiiiii252siQ"
_V3 = "iiiiii256sQ"
_KEYS4 = ("dimX", "dimY", "dimZ", "gridType", "elementType", "bytesPerElement", "info", "dimT", "timestamp")


class UniError(IOError):
    pass


def read_uni(path):
    """-> (header dict in the reference's key order, ndarray [Z,Y,X,C] (or [T,Z,Y,X,C] when dimT > 1))."""
    with gzip.open(path, "rb") as fh:
        ident = fh.read(4)
        raw = fh.read(288)
        if ident == b"MNT3":
            head = dict(zip(_KEYS4, struct.unpack(_V4, raw)))
        elif ident == b"MNT2":
            dx, dy, dz, gt, et, bpe, info, ts = struct.unpack(_V3, raw)
            head = dict(zip(_KEYS4, (dx, dy, dz, gt, et, bpe, info[:252], 0, ts)))
        else:
            raise UniError("%s: unsupported .uni id %r (4-D grids M4T2/M4T3 are not supported by the reference either)"
                           % (path, ident))
        et, bpe = head["elementType"], head["bytesPerElement"]
        if not ((bpe == 12 and et == 2) or (bpe == 4 and et in (0, 1))):
            raise UniError("%s: unsupported element type %d / %d bytes" % (path, et, bpe))
        data = np.frombuffer(fh.read(), dtype="int32" if et == 0 else "float32")
    ch = 3 if et == 2 else 1
    dims = [head["dimZ"], head["dimY"], head["dimX"], ch]
    if head["dimT"] > 1:
        dims = [head["dimT"]] + dims
    return head, data.reshape(dims)


def write_uni(path, head, content):
    """Always writes the v4 (`MNT3`) header and float32 content, like the reference writer."""
    vals = [head[k] for k in _KEYS4]
    content = np.ascontiguousarray(content, dtype=np.float32)
    n = head["dimX"] * head["dimY"] * head["dimZ"] * (3 if head["elementType"] == 2 else 1)
    if content.size != n:
        raise UniError("content has %d values, header describes %d" % (content.size, n))
    with gzip.open(path, "wb") as fh:
        fh.write(b"MNT3")
        fh.write(struct.pack(_V4, *vals))
        fh.write(memoryview(content.reshape(-1)))


def make_header(dim, element_type=1, timestamp=0):
    return dict(dimX=int(dim[2]), dimY=int(dim[1]), dimZ=int(dim[0]), gridType=1 if element_type == 1 else 4,
                elementType=int(element_type), bytesPerElement=12 if element_type == 2 else 4, info=b"\0" * 252, dimT=0,
                timestamp=int(timestamp))
