"""Host-side .uni output throughput: the reference-identical single gzip stream vs the multi-member
writer, on a synthetic thresholded density volume. CPU only.   python tools/bench_uni_io.py [S] [threads]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import mpgan_b200  # noqa: F401
from mpgan_b200 import io_pipeline, synth, uni

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
low = synth.synthetic_volume(S // 4, seed=1)[..., 0]
vol = np.kron(low, np.ones((4, 4, 4), np.float32)).astype(np.float32)
vol += 0.01 * np.random.default_rng(0).standard_normal(vol.shape).astype(np.float32) * (vol > 0)
vol[vol < 0.0005] = 0.0
head = uni.make_header((S, S, S), 1)
mb = vol.nbytes / 1e6
out = dict(volume="%d^3 float32 (%.0f MB), %.0f %% zeros" % (S, mb, 100 * float((vol == 0).mean())), threads=threads)
with tempfile.TemporaryDirectory() as d:
    t0 = time.time()
    uni.write_uni(os.path.join(d, "a.uni"), head, vol)
    t1 = time.time()
    io_pipeline.write_uni_parallel(os.path.join(d, "b.uni"), head, vol, threads=threads)
    t2 = time.time()
    _, back = uni.read_uni(os.path.join(d, "b.uni"))
    t3 = time.time()
    assert np.array_equal(back[..., 0], vol)
    out.update(reference_writer_s=t1 - t0, reference_writer_MBps=mb / (t1 - t0), size_a_MB=os.path.getsize(os.path.join(d, "a.uni")) / 1e6,
               parallel_writer_s=t2 - t1, parallel_writer_MBps=mb / (t2 - t1), size_b_MB=os.path.getsize(os.path.join(d, "b.uni")) / 1e6,
               read_back_s=t3 - t2)
print(json.dumps(out))
