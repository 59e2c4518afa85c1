"""One C4-shaped frame at 66.39 nm pixels (no box tables: every footprint gathers SAT corners), a few times, for
`ncu -k regex:render_strips`."""
import ctypes
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from microbench import engine_for  # noqa: E402
from scopyon_b200 import _native  # noqa: E402

size, n = 2048, 100000
yaml = """
default:
    magnification: 241
    light_source: {angle: {value: 0.0, units: radian}}
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: 16.0e-6, units: m}, QE: 0.73, exposure_time: 0.033}
""" % (size, size)
configs, eng = engine_for(yaml)
pl = configs.pixel_length
rng = numpy.random.RandomState(1)
data = numpy.zeros((n, 5))
data[:, 1:3] = rng.uniform(-size * pl / 2, size * pl / 2, (n, 2))
data[:, 0] = rng.uniform(0, 1.5e-6, n)
eng.ensure_all_tables()
soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).cuda()
w = torch.full((n,), 30.0, dtype=torch.float64, device="cuda")
out = torch.empty((size, size), dtype=torch.float32, device="cuda")
work = eng._render_workspace(n)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    eng._call("scb_render_expected", ctypes.byref(eng.geom), n, _native.ptr(soa[0]), _native.ptr(soa[1]),
              _native.ptr(soa[2]), _native.ptr(w), _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type,
              _native.ptr(eng.inv_scale), _native.ptr(eng.slot_of_key), _native.ptr(out), _native.F32, 0,
              _native.ptr(work), work.numel(), _native.ptr(eng.errors), eng._stream())
torch.cuda.synchronize()
print("done", float(out.double().sum()))
