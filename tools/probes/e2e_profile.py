"""Host-side profile (cProfile) of the end-to-end frame loop: generate_images on the C4 workload."""
import cProfile
import os
import pstats
import sys
import warnings

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import scopyon_b200  # noqa: E402
from bench import box, make_config  # noqa: E402

size, n, frames = 2048, 100000, 64
config = make_config(size)
lo, hi = box(size)
rng = numpy.random.RandomState(1)
base = numpy.empty((n, 5))
for k in range(3):
    base[:, k] = rng.uniform(lo[k], hi[k], n)
base[:, 3] = numpy.arange(n)
base[:, 4] = 1.0
inputs = [(k * 0.033, base + numpy.array([1e-9 * k, 0, 0, 0, 0])) for k in range(frames + 6)]
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(3))
    gen = sim.generate_images(inputs, num_frames=frames + 6)
    for _ in range(6):
        next(gen)
    torch.cuda.synchronize()
    prof = cProfile.Profile()
    import time
    t0 = time.perf_counter()
    prof.enable()
    total = 0.0
    for img in gen:
        total += float(img.as_array(numpy.float32)[0, 0])
    prof.disable()
    dt = time.perf_counter() - t0
print("ms per frame (under cProfile): %.3f" % (dt / frames * 1e3))
stats = pstats.Stats(prof)
stats.sort_stats("tottime").print_stats(28)
stats.sort_stats("cumulative").print_stats(30)
