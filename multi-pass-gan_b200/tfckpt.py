"""TensorFlow "tensor bundle" (checkpoint V2) reader / writer without TensorFlow.

The reference can only obtain weights through `tf.train.Saver.restore` (GAN/multipassGAN-out.py:367-386:
`model_%04d.ckpt.index` + `model_%04d.ckpt.data-00000-of-00001`, keys = variable names with the `gen_N/` scope and the
`:0` suffix stripped). TensorFlow is not installable in this image, so the format is restated here from its published
definition (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/{table,block,format}: a LevelDB-style sorted
string table whose values are BundleEntryProto messages):

  <prefix>.index                    SSTable: data blocks | metaindex block | index block | 48-byte footer
      block      = entries (shared|non_shared|value_len varint32, key delta, value) + uint32 restart offsets + count,
                   followed by a 1-byte compression type (0 none, 1 snappy) and a masked CRC32C of block + type
      footer     = metaindex BlockHandle, index BlockHandle (varint64 offset, size), zero padded to 40 bytes,
                   magic 0xdb4775248b80fb57 (little endian)
      key ""     -> BundleHeaderProto {num_shards=1, endianness=2 (0 = little), version=3}
      key <name> -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32), slices=7}
  <prefix>.data-00000-of-00001      raw little-endian tensor bytes, addressed by (offset, size)

PARITY UNPINNED: no TensorFlow-written checkpoint exists in the reference tree or in this image, so the reader is
tested against this module's own writer plus known-answer vectors of the primitives (CRC32C, masking, varints,
footer magic). `read_checkpoint` verifies block checksums, so a format misunderstanding fails loudly instead of
yielding wrong weights. Pinned by a third party where one exists in the image: the snappy decoder (compression type 1 of
the table format) decodes streams produced by Google's snappy library through pyarrow, also as the compressor of whole
index blocks (tests/test_tfckpt.py); multi-shard bundles (`num_shards` > 1, `shard_id`) round-trip through the writer.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
FOOTER_LEN = 48
BLOCK_TRAILER = 5
MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
           6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 19: np.dtype("<f2"),
           22: np.dtype("<u4"), 23: np.dtype("<u8")}
_DTYPE_CODE = {v: k for k, v in _DTYPES.items()}
DT_BFLOAT16 = 14


class CheckpointError(Exception):
    pass


# ------------------------------------------------------------------------------------------ CRC32C (Castagnoli)
def _make_crc_table():
    tab = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TABLE = _make_crc_table()
_CRC_NP = None


def crc32c(data, crc=0):
    """CRC-32C of bytes-like `data` (reflected polynomial 0x82F63B78)."""
    data = memoryview(data).cast("B")
    n = len(data)
    if n >= 1 << 16:
        return _crc32c_numpy(np.frombuffer(data, dtype=np.uint8), crc)
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _crc32c_numpy(buf, crc=0):
    """Large buffers: 4096 interleaved lanes advanced byte-by-byte with vectorised table look-ups, then combined
    lane by lane with the GF(2) 'append zeros' operator (same trick as zlib's crc32_combine)."""
    global _CRC_NP
    if _CRC_NP is None:
        _CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)
    tab = _CRC_NP
    lanes = 4096
    n = buf.size
    per = n // lanes
    head = buf[: per * lanes].reshape(lanes, per)
    c = np.full(lanes, 0, dtype=np.uint32)
    c[0] = np.uint32(crc ^ 0xFFFFFFFF)  # only the first lane carries the initial state
    for i in range(per):
        c = tab[(c ^ head[:, i]) & 0xFF] ^ (c >> np.uint32(8))
    # combine: crc(A||B) state = shift(state_A, len(B)) xor state_B(with zero init)
    op = _zeros_operator(per)
    total = int(c[0])
    for lane in range(1, lanes):
        total = _gf2_apply(op, total) ^ int(c[lane])
    cc = total
    tabl = _CRC_TABLE
    for b in buf[per * lanes:].tobytes():
        cc = tabl[(cc ^ b) & 0xFF] ^ (cc >> 8)
    return cc ^ 0xFFFFFFFF


def _gf2_apply(mat, vec):
    s = 0
    i = 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_apply(mat, mat[i]) for i in range(32)]


def _zeros_operator(nbytes):
    """32x32 GF(2) matrix advancing a raw CRC register over `nbytes` zero bytes."""
    one_bit = [0x82F63B78] + [1 << (i - 1) for i in range(1, 32)]  # one zero BIT
    m = one_bit
    for _ in range(3):  # -> one zero byte
        m = _gf2_square(m)
    result = [1 << i for i in range(32)]  # identity
    power = m
    k = nbytes
    while k:
        if k & 1:
            result = [_gf2_apply(power, result[i]) for i in range(32)]
        power = _gf2_square(power)
        k >>= 1
    return result


def mask_crc(crc):
    """leveldb/TF crc masking: rotate right by 15 and add a constant."""
    return (((crc >> 15) | (crc << 17)) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------ varints / protobuf wire
def put_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(buf, pos):
    shift = 0
    val = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _pb_fields(buf):
    """Yield (field_number, wire_type, value) of a serialized protobuf message."""
    pos = 0
    while pos < len(buf):
        key, pos = get_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def _pb_varint_field(fn, v):
    return put_varint(fn << 3) + put_varint(v)


def _pb_bytes_field(fn, b):
    return put_varint((fn << 3) | 2) + put_varint(len(b)) + b


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf):
    dims = []
    for fn, _, v in _pb_fields(buf):
        if fn == 2:  # Dim
            size = 0
            for f2, _, v2 in _pb_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif fn == 3 and v:
            raise CheckpointError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, slices=0)
    for fn, _, v in _pb_fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            e["shape"] = _parse_shape(v)
        elif fn == 3:
            e["shard_id"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc32c"] = v
        elif fn == 7:
            e["slices"] += 1
    return e


def _parse_header(buf):
    h = dict(num_shards=0, endianness=0, version=None)
    for fn, _, v in _pb_fields(buf):
        if fn == 1:
            h["num_shards"] = v
        elif fn == 2:
            h["endianness"] = v
        elif fn == 3:
            h["version"] = {f: x for f, _, x in _pb_fields(v)}
    return h


# ------------------------------------------------------------------------------------------ snappy (raw format) decoder
def snappy_decompress(data):
    n, pos = get_varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(data[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += data[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | data[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = data[pos] | (data[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(data[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):  # may overlap its own output
            out.append(out[-off])
    if len(out) != n:
        raise CheckpointError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------ SSTable
def _read_block(buf, offset, size, verify=True):
    raw = buf[offset:offset + size]
    if len(raw) != size or offset + size + BLOCK_TRAILER > len(buf):
        raise CheckpointError("block handle (%d, %d) outside the index file" % (offset, size))
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify:
        actual = crc32c(bytes(raw) + bytes([ctype]))
        if unmask_crc(stored) != actual:
            raise CheckpointError("index block checksum mismatch at offset %d" % offset)
    if ctype == 0:
        return bytes(raw)
    if ctype == 1:
        return snappy_decompress(bytes(raw))
    raise CheckpointError("unknown block compression type %d" % ctype)


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("block too small")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * nrestart
    if limit < 0:
        raise CheckpointError("corrupt restart array")
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = get_varint(block, pos)
        non_shared, pos = get_varint(block, pos)
        vlen, pos = get_varint(block, pos)
        if shared > len(key):
            raise CheckpointError("corrupt key prefix")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _table_items(buf, verify=True):
    if len(buf) < FOOTER_LEN:
        raise CheckpointError("index file shorter than a table footer")
    footer = buf[-FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError("not an SSTable (bad magic): is this a V1 checkpoint?")
    _, p = get_varint(footer, 0)      # metaindex handle (unused)
    _, p = get_varint(footer, p)
    ioff, p = get_varint(footer, p)
    isize, p = get_varint(footer, p)
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, q = get_varint(handle, 0)
        size, q = get_varint(handle, q)
        for key, value in _block_entries(_read_block(buf, off, size, verify)):
            yield key, value


# ------------------------------------------------------------------------------------------ public API
def list_checkpoint(prefix):
    """name -> dict(dtype, shape, offset, size, ...) of every tensor in the bundle `prefix` (+ '' -> header)."""
    with open(prefix + ".index", "rb") as fh:
        buf = fh.read()
    out = {}
    for key, value in _table_items(buf):
        name = key.decode("utf-8")
        out[name] = _parse_header(value) if name == "" else _parse_entry(value)
    if "" not in out:
        raise CheckpointError("bundle header missing")
    if out[""]["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return out


def read_checkpoint(prefix, names=None, verify_data=False):
    """Read tensors of a V2 checkpoint: {name: numpy array}. `names`: iterable of wanted tensors (default: all numeric
    ones). `verify_data`: also check each tensor's CRC32C (the index blocks are always verified)."""
    entries = list_checkpoint(prefix)
    header = entries.pop("")
    nshards = max(1, header["num_shards"])
    wanted = sorted(entries) if names is None else list(names)
    shards = {}
    out = {}
    for name in wanted:
        if name not in entries:
            raise KeyError("tensor '%s' is not in checkpoint %s (has: %s ...)" % (name, prefix, ", ".join(sorted(entries)[:5])))
        e = entries[name]
        if e["slices"]:
            raise CheckpointError("partitioned variable '%s' is not supported" % name)
        if e["dtype"] == DT_BFLOAT16:
            dt = np.dtype("<u2")
        elif e["dtype"] in _DTYPES:
            dt = _DTYPES[e["dtype"]]
        elif names is None:
            continue  # strings / resources: skipped when reading everything
        else:
            raise CheckpointError("tensor '%s' has unsupported dtype %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, nshards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if raw.size != e["size"] or count * dt.itemsize != e["size"]:
            raise CheckpointError("tensor '%s': %d bytes on disk, shape %s needs %d" % (name, raw.size, e["shape"], count * dt.itemsize))
        if verify_data and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise CheckpointError("tensor '%s': data checksum mismatch" % name)
        arr = np.frombuffer(bytes(raw), dtype=dt).reshape(e["shape"])
        if e["dtype"] == DT_BFLOAT16:
            arr = (arr.astype(np.uint32) << 16).view(np.float32)
        out[name] = arr
    return out


def _build_block(items, restart_interval=16):
    out = bytearray()
    restarts = []
    last = b""
    for i, (key, value) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        out += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts.append(0)
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_checkpoint(prefix, tensors, block_size=4096, compressor=None, num_shards=1):
    """Write {name: array} as a V2 checkpoint (`prefix`.index + `prefix`.data-0000k-of-0000N) in the layout of
    tf.train.Saver. CRCs on every block and tensor. Defaults: one data shard, uncompressed index blocks (what TF writes).
    compressor: callable bytes -> raw snappy stream; index blocks are then stored with compression type 1 (the LevelDB table
    format's kSnappyCompression) -- used by the tests with a third-party snappy implementation to pin the reader's decoder.
    num_shards > 1: tensors are dealt round-robin over that many data files (shard_id / per-shard offsets in the entries),
    the layout a sharded Saver merges its bundles into."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    if "" in tensors:
        raise ValueError("the empty name is reserved for the bundle header")
    num_shards = int(num_shards)
    if num_shards < 1:
        raise ValueError("num_shards must be >= 1")
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    if num_shards > 1:
        return _write_sharded(prefix, tensors, names, block_size, compressor, num_shards)
    entries = []
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in names:
            a = np.asarray(tensors[name])  # (ascontiguousarray would turn a scalar into shape (1,))
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _DTYPE_CODE:
                raise ValueError("dtype %s of '%s' is not supported" % (a.dtype, name))
            raw = a.astype(dt, copy=False).tobytes(order="C")
            fh.write(raw)
            shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in a.shape)
            msg = _pb_varint_field(1, _DTYPE_CODE[np.dtype(dt)]) + _pb_bytes_field(2, shape)
            if offset:
                msg += _pb_varint_field(4, offset)
            msg += _pb_varint_field(5, len(raw)) + put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc32c(raw)))
            entries.append((name.encode("utf-8"), msg))
            offset += len(raw)
    _write_index(prefix, entries, 1, block_size, compressor)


def _entry_message(a, dt, raw, offset, shard_id):
    shape = b"".join(_pb_bytes_field(2, _pb_varint_field(1, int(d))) for d in a.shape)
    msg = _pb_varint_field(1, _DTYPE_CODE[np.dtype(dt)]) + _pb_bytes_field(2, shape)
    if shard_id:
        msg += _pb_varint_field(3, shard_id)
    if offset:
        msg += _pb_varint_field(4, offset)
    return msg + _pb_varint_field(5, len(raw)) + put_varint((6 << 3) | 5) + struct.pack("<I", mask_crc(crc32c(raw)))


def _write_sharded(prefix, tensors, names, block_size, compressor, num_shards):
    files = [open("%s.data-%05d-of-%05d" % (prefix, k, num_shards), "wb") for k in range(num_shards)]
    offsets = [0] * num_shards
    entries = []
    try:
        for i, name in enumerate(names):
            a = np.asarray(tensors[name])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in _DTYPE_CODE:
                raise ValueError("dtype %s of '%s' is not supported" % (a.dtype, name))
            raw = a.astype(dt, copy=False).tobytes(order="C")
            k = i % num_shards
            files[k].write(raw)
            entries.append((name.encode("utf-8"), _entry_message(a, dt, raw, offsets[k], k)))
            offsets[k] += len(raw)
    finally:
        for fh in files:
            fh.close()
    _write_index(prefix, entries, num_shards, block_size, compressor)


def _write_index(prefix, entries, num_shards, block_size, compressor):
    header = _pb_varint_field(1, num_shards) + _pb_bytes_field(3, _pb_varint_field(1, 1))  # num_shards, little endian, producer 1
    items = [(b"", header)] + entries
    blocks = []
    cur, cur_bytes = [], 0
    for it in items:
        cur.append(it)
        cur_bytes += len(it[0]) + len(it[1]) + 8
        if cur_bytes >= block_size:
            blocks.append(cur)
            cur, cur_bytes = [], 0
    if cur:
        blocks.append(cur)
    out = bytearray()

    def emit(block_bytes):
        off = len(out)
        ctype = b"\x00"  # kNoCompression
        if compressor is not None:
            block_bytes, ctype = bytes(compressor(block_bytes)), b"\x01"  # kSnappyCompression
        out.extend(block_bytes)
        out.extend(ctype)
        out.extend(struct.pack("<I", mask_crc(crc32c(block_bytes + ctype))))
        return put_varint(off) + put_varint(len(block_bytes))

    index_items = []
    for blk in blocks:
        handle = emit(_build_block(blk))
        index_items.append((blk[-1][0], handle))  # separator key = last key of the block
    meta_handle = emit(_build_block([]))
    index_handle = emit(_build_block(index_items, restart_interval=1))
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as fh:
        fh.write(bytes(out))


def load_generator_weights(prefix, graph_names, scope):
    """The reference's restore (GAN/multipassGAN-out.py:367-386): the Saver is built with
    {var.name[6:-2]: var}, i.e. checkpoint key = graph name without the `gen_N/` scope. Returns
    {graph name: array} for every name in `graph_names` (all must start with `scope` + '/')."""
    pre = scope + "/"
    keys = {}
    for n in graph_names:
        if not n.startswith(pre):
            raise ValueError("variable '%s' is outside scope '%s'" % (n, scope))
        keys[n] = n[len(pre):]
    got = read_checkpoint(prefix, names=sorted(set(keys.values())), verify_data=True)  # CRC32C of every tensor: corrupt .data fails loudly
    return {n: np.asarray(got[k], dtype=np.float32) for n, k in keys.items()}
