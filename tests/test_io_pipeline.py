"""Frame-loop I/O pipeline: ordering, overlap, error propagation, multi-member .uni writer."""
import threading
import time

import numpy as np
import pytest

import mpgan_b200  # noqa: F401
from mpgan_b200 import io_pipeline, uni


def test_parallel_uni_writer_reads_back_identically(tmp_path):
    rng = np.random.default_rng(0)
    vol = rng.random((24, 20, 16, 1), dtype=np.float32)
    vol[vol < 0.5] = 0.0
    head = uni.make_header((24, 20, 16), 1, timestamp=7)
    a, b = str(tmp_path / "a.uni"), str(tmp_path / "b.uni")
    uni.write_uni(a, head, vol)
    io_pipeline.write_uni_parallel(b, head, vol, threads=3, chunk_bytes=4096)  # 8 gzip members
    ha, va = uni.read_uni(a)
    hb, vb = uni.read_uni(b)
    assert ha == hb
    np.testing.assert_array_equal(va, vb)
    np.testing.assert_array_equal(vb, vol)
    with pytest.raises(uni.UniError):
        io_pipeline.write_uni_parallel(b, head, vol[:3])


def test_pipeline_orders_overlaps_and_stores_everything():
    log, stored = [], {}
    lock = threading.Lock()

    def load(f):
        time.sleep(0.05)
        return np.full(4, f, np.float32)

    def compute(f, x):
        with lock:
            log.append(f)
        time.sleep(0.02)
        return x * 2

    def store(f, y):
        time.sleep(0.08)
        with lock:
            stored[f] = y.copy()

    t0 = time.time()
    stats = io_pipeline.FramePipeline(load, compute, store, prefetch=2, readers=2, writers=4).run(range(3, 11))
    wall = time.time() - t0
    assert log == list(range(3, 11))                       # compute sees frames in order
    assert sorted(stored) == list(range(3, 11)) and all(float(stored[f][0]) == 2 * f for f in stored)
    assert stats["frames"] == 8 and stats["store_s"] >= 8 * 0.08 * 0.9
    serial = 8 * (0.05 + 0.02 + 0.08)
    assert wall < 0.7 * serial, (wall, serial)             # the three stages overlapped


def test_pipeline_propagates_errors():
    def bad_load(f):
        if f == 2:
            raise IOError("frame 2 is missing")
        return f

    with pytest.raises(IOError):
        io_pipeline.FramePipeline(bad_load, lambda f, x: x, lambda f, y: None).run(range(5))

    def bad_store(f, y):
        raise ValueError("disk full")

    with pytest.raises(ValueError):
        io_pipeline.FramePipeline(lambda f: f, lambda f, x: x, bad_store, writers=2).run(range(4))

    def bad_compute(f, x):
        raise RuntimeError("kernel failed")

    with pytest.raises(RuntimeError):
        io_pipeline.FramePipeline(lambda f: f, bad_compute, lambda f, y: None).run(range(3))
