"""End-to-end frame rate of generate_images (C4 workload) against the number of host widening threads."""
import os
import sys
import time
import warnings

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import scopyon_b200  # noqa: E402
from bench import box, make_config  # noqa: E402
from scopyon_b200 import _native  # noqa: E402

lib = _native.load()
size, n, frames = 2048, 100000, 96
config = make_config(size)
lo, hi = box(size)
rng = numpy.random.RandomState(1)
base = numpy.empty((n, 5))
for k in range(3):
    base[:, k] = rng.uniform(lo[k], hi[k], n)
base[:, 3] = numpy.arange(n)
base[:, 4] = 1.0
inputs = [(k * 0.033, base + numpy.array([1e-9 * k, 0, 0, 0, 0])) for k in range(frames + 6)]
for threads in [int(v) for v in sys.argv[1:]] or [2, 4, 6, 8, 12]:
    lib.scb_host_widen_threads(threads)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(3))
        gen = sim.generate_images(inputs, num_frames=frames + 6)
        for _ in range(6):
            next(gen)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        total = 0.0
        for img in gen:
            total += float(img.as_array()[0, 0])
        dt = time.perf_counter() - t0
    print("widen threads %2d: %.3f ms per frame, %.0f frames/s" % (threads, dt / frames * 1e3, frames / dt), flush=True)

# the float32 download alone (pinned, 2048 x 2048)
dev = torch.empty((size, size), dtype=torch.float32, device="cuda")
host = torch.empty((size, size), dtype=torch.float32, pin_memory=True)
for dtype in (torch.float32, torch.float64):
    dev = torch.empty((size, size), dtype=dtype, device="cuda")
    host = torch.empty((size, size), dtype=dtype, pin_memory=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        host.copy_(dev, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    print("D2H %s: %.3f ms per frame" % (dtype, a.elapsed_time(b) / 20))
