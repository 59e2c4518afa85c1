"""Repeat one kernel scenario a few times (for `ncu -k regex:... -s N -c M`).
  python tools/profile_driver.py detector|render3d|render2d [iterations]"""
import ctypes
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from microbench import engine_for  # noqa: E402
from bench import C4_YAML  # noqa: E402
from scopyon_b200 import _native  # noqa: E402

what = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if what == "detector":
    size = 4096
    configs, eng = engine_for(C4_YAML % (size, size, 2.5))
    photons = torch.full((size, size), 0.6, dtype=torch.float32, device="cuda")
    adc = torch.empty_like(photons)
    for k in range(iters):
        eng.detect(photons, k, 42, adc=adc)
else:
    size, n = 2048, 100000
    configs, eng = engine_for(C4_YAML % (size, size, 2.5))
    pl = configs.pixel_length
    rng = numpy.random.RandomState(1)
    data = numpy.zeros((n, 5))
    data[:, 1:3] = rng.uniform(-size * pl / 2, size * pl / 2, (n, 2))
    if what == "render3d":
        data[:, 0] = rng.uniform(0, 1.5e-6, n)
        eng.ensure_all_tables()
    else:
        eng.ensure_tables([0])
    soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).cuda()
    w = torch.full((n,), 30.0, dtype=torch.float64, device="cuda")
    out = torch.empty((size, size), dtype=torch.float32, device="cuda")
    work = eng._render_workspace(n)
    for k in range(iters):
        eng._call("scb_render_expected", ctypes.byref(eng.geom), n, _native.ptr(soa[0]), _native.ptr(soa[1]),
                  _native.ptr(soa[2]), _native.ptr(w), _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type, _native.ptr(eng.inv_scale),
                  _native.ptr(eng.slot_of_key), _native.ptr(out), _native.F32, 0, _native.ptr(work), work.numel(),
                  _native.ptr(eng.errors), eng._stream())
torch.cuda.synchronize()
print("done", what, iters)
