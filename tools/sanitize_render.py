"""Small render + detector workload for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_render.py"""
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from microbench import engine_for  # noqa: E402

YAML = """
default:
    magnification: 100
    detector: {type: CMOS, image_size: [200, 136], pixel_length: {value: %g, units: m}, QE: 0.73}
"""
for pixel, precision in ((6.5e-6, "f32"), (6.5e-6, "f64"), (6.639e-6, "f32")):   # box-table path twice, gather path once
    configs, eng = engine_for(YAML % pixel, precision=precision)
    pl = configs.pixel_length
    rng = numpy.random.RandomState(3)
    n = 1500
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 1.2e-6, n)
    data[:, 1] = rng.uniform(-110 * pl, 110 * pl, n)
    data[:, 2] = rng.uniform(-75 * pl, 75 * pl, n)
    data[:50, 1:3] = numpy.round(data[:50, 1:3] / pl) * pl          # exact pixel centres: the irregular path
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    out = torch.empty((eng.n_w, eng.n_h), dtype=eng.dtype, device=eng.device)
    img, _ = eng.render_expected([(0.033, data)], out=out)
    adc = eng.detect(img, 0, 7)
    torch.cuda.synchronize()
    print(pixel, precision, float(img.sum()), float(adc.mean()), int(eng.errors.item()))
print("done")
