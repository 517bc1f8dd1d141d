"""Frame-loop I/O pipeline of the apply entry point.

The reference's frame loop (GAN/multipassGAN-out.py:629-632 -> generate3DUniForNewNetwork :390-618) is strictly serial:
gunzip the input grids, run the networks, gzip the 512^3 result (`uniio.writeUni`, tools_wscale/uniio.py:95-123, Python's
gzip at its default level 9), next frame. Once the networks take ~0.3 s per frame the two gzip stages are >95 % of the
wall time. zlib releases the GIL, so plain threads overlap them with the GPU work:

    reader thread(s)  --queue-->  main thread (H2D, kernels, D2H)  --queue-->  writer threads (gzip + write)

`FramePipeline` is that three-stage pipeline (bounded queues, exceptions re-raised in the caller, results stored in
frame order independent of completion order). `write_uni_parallel` additionally compresses ONE file on several threads
as concatenated gzip members (RFC 1952 section 2.2; Python's gzip module and zlib's gzread -- mantaflow's reader -- both
read multi-member files transparently). The single-stream writer stays the default because it is byte-identical to
the reference writer.
"""
import queue
import struct
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import uni


def write_uni_parallel(path, head, content, threads=8, chunk_bytes=8 << 20, level=6):
    """`uni.write_uni` with the payload deflated in `chunk_bytes` pieces on `threads` threads, each piece a complete gzip
    member. Same decompressed bytes as the reference writer, different (and ~`threads`x faster) container."""
    vals = [head[k] for k in uni._KEYS4]
    content = np.ascontiguousarray(content, dtype=np.float32)
    n = head["dimX"] * head["dimY"] * head["dimZ"] * (3 if head["elementType"] == 2 else 1)
    if content.size != n:
        raise uni.UniError("content has %d values, header describes %d" % (content.size, n))
    header = b"MNT3" + struct.pack(uni._V4, *vals)
    body = memoryview(content.reshape(-1)).cast("B")
    pieces = [(0, header)]
    for i, off in enumerate(range(0, len(body), chunk_bytes)):
        pieces.append((i + 1, body[off:off + chunk_bytes]))

    def deflate(piece):
        co = zlib.compressobj(level, zlib.DEFLATED, 31)  # wbits 31: gzip header + trailer around the raw deflate stream
        return co.compress(piece[1]) + co.flush()

    with ThreadPoolExecutor(max_workers=max(1, int(threads))) as pool:
        members = list(pool.map(deflate, pieces))
    with open(path, "wb") as fh:
        for m in members:
            fh.write(m)


class FramePipeline:
    """load(frame) -> x ; compute(frame, x) -> y ; store(frame, y). `load` runs on `readers` threads up to `prefetch`
    frames ahead, `compute` on the calling thread in frame order, `store` on `writers` threads."""

    def __init__(self, load, compute, store, prefetch=2, readers=1, writers=4):
        self.load, self.compute, self.store = load, compute, store
        self.prefetch, self.readers, self.writers = max(1, int(prefetch)), max(1, int(readers)), max(1, int(writers))

    def run(self, frames):
        frames = list(frames)
        stats = dict(frames=len(frames), load_s=0.0, compute_s=0.0, store_s=0.0, wait_input_s=0.0, wall_s=0.0)
        lock = threading.Lock()
        errors = []
        t_start = time.time()
        loaded = {}
        loaded_cv = threading.Condition()
        next_to_load = [0]
        consumed = [0]

        def reader():
            while True:
                with loaded_cv:
                    while not errors and next_to_load[0] < len(frames) and next_to_load[0] - consumed[0] >= self.prefetch:
                        loaded_cv.wait(0.05)
                    if errors or next_to_load[0] >= len(frames):
                        return
                    i = next_to_load[0]
                    next_to_load[0] += 1
                try:
                    t0 = time.time()
                    x = self.load(frames[i])
                    with lock:
                        stats["load_s"] += time.time() - t0
                except BaseException as e:  # noqa: BLE001 - re-raised in the caller
                    x = e
                with loaded_cv:
                    loaded[i] = x
                    loaded_cv.notify_all()

        out_q = queue.Queue(maxsize=self.writers + 1)

        def writer():
            while True:
                item = out_q.get()
                if item is None:
                    return
                f, y = item
                try:
                    t0 = time.time()
                    self.store(f, y)
                    with lock:
                        stats["store_s"] += time.time() - t0
                except BaseException as e:  # noqa: BLE001
                    errors.append(e)

        rthreads = [threading.Thread(target=reader, daemon=True) for _ in range(self.readers)]
        wthreads = [threading.Thread(target=writer, daemon=True) for _ in range(self.writers)]
        for t in rthreads + wthreads:
            t.start()
        try:
            for i, f in enumerate(frames):
                t0 = time.time()
                with loaded_cv:
                    while i not in loaded:
                        if errors:  # a writer failed: the readers have stopped, do not wait for them
                            raise errors[0]
                        loaded_cv.wait(0.05)
                    x = loaded.pop(i)
                    consumed[0] = i + 1
                    loaded_cv.notify_all()
                stats["wait_input_s"] += time.time() - t0
                if isinstance(x, BaseException):
                    raise x
                if errors:
                    raise errors[0]
                t0 = time.time()
                y = self.compute(f, x)
                stats["compute_s"] += time.time() - t0
                out_q.put((f, y))
        except BaseException as e:
            errors.append(e)
            with loaded_cv:
                loaded_cv.notify_all()
            raise
        finally:
            for _ in wthreads:
                out_q.put(None)
            for t in wthreads:
                t.join()
            stats["wall_s"] = time.time() - t_start
        if errors:
            raise errors[0]
        return stats
