"""TensorFlow checkpoint-V2 (tensor bundle) reader / writer: known-answer vectors of the primitives and
round trips through this module's own writer. No TensorFlow-written file exists in the reference tree, so parity with
TF itself is unpinned (stated in the module docstring)."""
import os
import struct

import numpy as np
import pytest

import mpgan_b200  # noqa: F401
from mpgan_b200 import cli, tfckpt


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors + the classic check value
    assert tfckpt.crc32c(b"123456789") == 0xE3069283
    assert tfckpt.crc32c(bytes(32)) == 0x8A9136AA
    assert tfckpt.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfckpt.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfckpt.crc32c(b"") == 0
    # incremental == one shot; the vectorised large-buffer path == the byte loop
    rng = np.random.default_rng(0)
    big = rng.integers(0, 256, size=(1 << 16) * 3 + 12345, dtype=np.uint8).tobytes()
    assert tfckpt.crc32c(big[1000:], tfckpt.crc32c(big[:1000])) == tfckpt.crc32c(big)
    slow = 0xFFFFFFFF
    for b in big:
        slow = tfckpt._CRC_TABLE[(slow ^ b) & 0xFF] ^ (slow >> 8)
    assert tfckpt.crc32c(big) == slow ^ 0xFFFFFFFF


def test_crc_masking_and_varints():
    crc = tfckpt.crc32c(b"foo")
    assert tfckpt.unmask_crc(tfckpt.mask_crc(crc)) == crc and tfckpt.mask_crc(crc) != crc
    assert tfckpt.mask_crc(0) == 0xA282EAD8  # rotate(0) + kMaskDelta
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 - 1):
        enc = tfckpt.put_varint(v)
        assert tfckpt.get_varint(enc, 0) == (v, len(enc))
    assert tfckpt.put_varint(300) == b"\xac\x02"
    assert len(tfckpt.put_varint(-1)) == 10  # negative int64 -> 10-byte two's complement varint


def test_snappy_decoder_literals_and_copies():
    # "abcdabcdabcdab": literal "abcd" + copy(offset 4, length 10) with a 1-byte offset tag
    stream = tfckpt.put_varint(14) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((10 - 4) << 2) | 1, 4])
    assert tfckpt.snappy_decompress(stream) == b"abcdabcdabcdab"
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.snappy_decompress(tfckpt.put_varint(4) + bytes([((4 - 4) << 2) | 1, 9]))


def _tensors(rng, n):
    out = {}
    for i in range(n):
        shape = tuple(int(s) for s in rng.integers(1, 6, size=int(rng.integers(0, 5))))
        out["generator/genBlock%d/g_c%s_%d/%s" % (2 ** (i % 3 + 1), "AB"[i % 2], i, ("weight", "bias")[i % 2])] = \
            rng.standard_normal(shape).astype(np.float32)
    out["global_step"] = np.array(1234, dtype=np.int64)
    out["generator/wide/weight"] = rng.standard_normal((5, 5, 96, 96)).astype(np.float32)  # ~0.9 MB: vectorised CRC path
    out["half"] = rng.standard_normal((3, 7)).astype(np.float16)
    return out


def test_write_read_roundtrip_many_blocks(tmp_path):
    rng = np.random.default_rng(1)
    tensors = _tensors(rng, 150)  # > 4 KB of entries: several data blocks, prefix-compressed keys
    prefix = str(tmp_path / "test_0001" / "model_0002.ckpt")
    tfckpt.write_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    with open(prefix + ".index", "rb") as fh:
        raw = fh.read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57
    listing = tfckpt.list_checkpoint(prefix)
    assert listing[""]["num_shards"] == 1 and listing[""]["endianness"] == 0
    assert set(listing) - {""} == set(tensors)
    got = tfckpt.read_checkpoint(prefix, verify_data=True)
    for name, a in tensors.items():
        assert got[name].dtype == a.dtype and got[name].shape == a.shape
        np.testing.assert_array_equal(got[name], a)
    sub = tfckpt.read_checkpoint(prefix, names=["global_step"])
    assert list(sub) == ["global_step"] and int(sub["global_step"]) == 1234
    with pytest.raises(KeyError):
        tfckpt.read_checkpoint(prefix, names=["nope"])


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(2)
    prefix = str(tmp_path / "m.ckpt")
    tfckpt.write_checkpoint(prefix, {"a/weight": rng.standard_normal((4, 4)).astype(np.float32)})
    raw = bytearray(open(prefix + ".index", "rb").read())
    bad = bytearray(raw)
    bad[3] ^= 0x40  # inside the first data block
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.read_checkpoint(prefix)
    bad = bytearray(raw)
    bad[-1] ^= 0xFF  # footer magic
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.read_checkpoint(prefix)
    open(prefix + ".index", "wb").write(bytes(raw))
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    tfckpt.read_checkpoint(prefix)  # data CRC is opt-in ...
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.read_checkpoint(prefix, verify_data=True)  # ... and catches the flipped bit


def test_generator_scope_stripping_matches_reference_saver(tmp_path):
    """GAN/multipassGAN-out.py:367-371: Saver(var_list={var.name[6:-2]: var}) -> keys lose the `gen_N/` scope."""
    from mpgan_b200 import pipeline as P
    w = P.make_weights_out(8, 5, upRes=8, nets=(1, 2))
    for i in (1, 2):
        prefix = str(tmp_path / ("test_%04d" % i) / "model_0007.ckpt")
        tfckpt.write_checkpoint(prefix, {k[len("gen_%d/" % i):]: v for k, v in w[i].items()})
        assert all(k.startswith("generator/") for k in tfckpt.list_checkpoint(prefix) if k)
        got = tfckpt.load_generator_weights(prefix, sorted(w[i]), "gen_%d" % i)
        assert set(got) == set(w[i])
        for k in w[i]:
            np.testing.assert_array_equal(got[k], w[i][k])
    with pytest.raises(ValueError):
        tfckpt.load_generator_weights(prefix, ["gen_1/generator/x"], "gen_2")


def test_cli_reports_missing_checkpoint(tmp_path):
    with pytest.raises(SystemExit) as e:
        cli.main(["prog", "basePath", str(tmp_path) + "/", "load_model_test_1", "3", "load_model_no_1", "9", "useVelocities", "1",
                  "simSize", "4", "tileSize", "4"])
    assert "model_0009.ckpt" in str(e.value)


@pytest.mark.gpu
def test_cli_restores_checkpoints_like_the_reference(tmp_path):
    """End to end: weights written as TF checkpoints under basePath/test_%04d/ are restored by the flags the reference
    uses (load_model_test_N / load_model_no_N) and give exactly the frames of the same weights passed in directly."""
    from mpgan_b200 import pipeline as P, synth, uni
    L, u = 4, 4
    sim = tmp_path / "sim_1000"
    sim.mkdir()
    x = synth.synthetic_volume(L, seed=4)
    uni.write_uni(str(sim / "density_low_0000.uni"), uni.make_header((L, L, L), 1), x[..., 0:1])
    uni.write_uni(str(sim / "velocity_low_0000.uni"), uni.make_header((L, L, L), 2), x[..., 1:4])
    specs = {1: P.NetSpec(True, True, 32, 32, 3, True), 2: P.NetSpec(True, False, 32, 32, 5)}
    w = P.make_weights_out(L, 9, upRes=u, specs=specs, nets=(1, 2))
    for i, (test_no, model_no) in ((1, (11, 3)), (2, (12, 4))):
        tfckpt.write_checkpoint(str(tmp_path / ("test_%04d" % test_no) / ("model_%04d.ckpt" % model_no)),
                                {k[len("gen_%d/" % i):]: v for k, v in w[i].items()})
    flags = dict(out=1, precision="fp32", basePath=str(tmp_path) + "/", packedSimPath=str(tmp_path) + "/", fromSim=1000,
                 frame_min=0, frame_max=1, simSize=L, tileSize=L, upRes=u, useVelocities=1, genUni=1, transposeAxis=0,
                 pixelNorm=1, batchNorm=0, addBicubicUpsample=1, upsampleMode=1, firstNNArch=1, velScale=1.0,
                 load_model_test_1=11, load_model_no_1=3, use_res_net1=1, add_adj_idcs1=1, startFms1=32, maxFms1=32, filterSize1=3,
                 load_model_test_2=12, load_model_no_2=4, use_res_net2=1, add_adj_idcs2=0, startFms2=32, maxFms2=32, filterSize2=5)
    argv = ["multipassGAN-out.py"]
    for k, v in flags.items():
        argv += [k, str(v)]
    assert cli.main(argv) == 0
    mp = P.MultiPassOut(L, w, upRes=u, specs=specs, precision="fp32")
    head, vol = uni.read_uni(str(sim / "source_0000.uni"))
    np.testing.assert_array_equal(vol[..., 0], mp(x).cpu().numpy())


def test_convert_tool_roundtrip(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("ckpt_convert", os.path.join(os.path.dirname(__file__), "..", "tools", "ckpt_convert.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    rng = np.random.default_rng(6)
    w = {"gen_1/generator/a/weight": rng.standard_normal((3, 3, 4, 8)).astype(np.float32),
         "gen_1/generator/a/bias": np.full(8, 0.1, np.float32), "gen_2/generator/b/weight": np.ones((1, 1, 2, 2), np.float32)}
    npz = str(tmp_path / "w.npz")
    np.savez(npz, **w)
    prefix = str(tmp_path / "test_0001" / "model_0001.ckpt")
    assert tool.main(["to-ckpt", npz, prefix, "--strip", "gen_1"]) == 0
    assert sorted(k for k in tfckpt.list_checkpoint(prefix) if k) == ["generator/a/bias", "generator/a/weight"]
    back = str(tmp_path / "back.npz")
    assert tool.main(["to-npz", prefix, back, "--scope", "gen_1"]) == 0
    got = np.load(back)
    assert sorted(got.files) == ["gen_1/generator/a/bias", "gen_1/generator/a/weight"]
    np.testing.assert_array_equal(got["gen_1/generator/a/weight"], w["gen_1/generator/a/weight"])
    assert tool.main(["list", prefix]) == 0


def _pa_snappy():
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("snappy"):
        pytest.skip("pyarrow built without snappy")
    return lambda b: pa.compress(bytes(b), codec="snappy", asbytes=True)


def test_snappy_decoder_against_a_third_party_compressor():
    """Streams produced by Google's snappy library (through pyarrow) -- literals of every length class, copies with 1-, 2-
    and 4-byte offsets, overlapping copies -- decode to the original bytes."""
    comp = _pa_snappy()
    rng = np.random.default_rng(12)
    cases = [b"", b"a", b"abcd" * 5, bytes(range(256)) * 3, rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),
             (b"the quick brown fox " * 40 + rng.integers(0, 4, 3000, dtype=np.uint8).tobytes()) * 30,
             b"\x00" * 200000, rng.integers(0, 2, 150000, dtype=np.uint8).tobytes()]
    big = rng.integers(0, 256, 90000, dtype=np.uint8).tobytes()
    cases.append(big + b"x" * 10 + big)            # a copy further back than 64 KiB
    for raw in cases:
        enc = comp(raw)
        assert tfckpt.snappy_decompress(enc) == raw
    assert sum(len(comp(c)) for c in cases) < sum(len(c) for c in cases) // 2   # the streams really are compressed


def test_index_blocks_compressed_by_a_third_party_snappy(tmp_path):
    """A bundle whose index blocks carry LevelDB's kSnappyCompression, compressed by Google's snappy (pyarrow)."""
    comp = _pa_snappy()
    rng = np.random.default_rng(13)
    tensors = _tensors(rng, 120)
    prefix = str(tmp_path / "model_0001.ckpt")
    tfckpt.write_checkpoint(prefix, tensors, block_size=1024, compressor=comp)
    plain = str(tmp_path / "plain.ckpt")
    tfckpt.write_checkpoint(plain, tensors, block_size=1024)
    assert os.path.getsize(prefix + ".index") < os.path.getsize(plain + ".index")
    got = tfckpt.read_checkpoint(prefix, verify_data=True)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    # a flipped byte inside a compressed block fails the block checksum
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[40] ^= 0x10
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.read_checkpoint(prefix)


def test_multi_shard_bundle(tmp_path):
    """num_shards = 3: .data-0000k-of-00003 files, shard_id / per-shard offsets in the entries (the layout a sharded Saver
    merges into); a missing shard and a corrupted one fail loudly."""
    rng = np.random.default_rng(14)
    tensors = _tensors(rng, 40)
    prefix = str(tmp_path / "model_0002.ckpt")
    tfckpt.write_checkpoint(prefix, tensors, num_shards=3)
    assert sorted(f for f in os.listdir(tmp_path) if ".data-" in f) == ["model_0002.ckpt.data-%05d-of-00003" % k for k in range(3)]
    got = tfckpt.read_checkpoint(prefix, verify_data=True)
    for k, v in tensors.items():
        assert np.array_equal(got[k], v) and got[k].dtype == v.dtype, k
    some = sorted(tensors)[4]
    assert np.array_equal(tfckpt.read_checkpoint(prefix, names=[some])[some], tensors[some])
    shard1 = prefix + ".data-00001-of-00003"
    data = bytearray(open(shard1, "rb").read())
    data[7] ^= 1
    open(shard1, "wb").write(bytes(data))
    with pytest.raises(tfckpt.CheckpointError):
        tfckpt.read_checkpoint(prefix, verify_data=True)
    os.remove(shard1)
    with pytest.raises((tfckpt.CheckpointError, OSError)):
        tfckpt.read_checkpoint(prefix, verify_data=True)
