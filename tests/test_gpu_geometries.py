"""Odd detector geometries through the engine, against the numpy oracle.

The render kernel has three size-dependent regimes: images narrower than one strip,
footprints wider than a strip (10 nm pixels: 200-pixel stamps, 208 slots per phase) and
footprints of one or two pixels (2 um pixels: 2 slots per phase).  Each must agree with the
reference's accumulation (scopyon/epifm.py:432-517, restated in oracle/epifm_oracle.py).
"""
import numpy
import pytest
import torch

from conftest import gpu_engine
import epifm_oracle as orc

pytestmark = pytest.mark.gpu

YAML = ("default:\n    magnification: %g\n    detector: {type: CMOS, image_size: [%d, %d], "
        "pixel_length: {value: %.9e, units: m}}\n")

CASES = [
    # name, magnification, n_w, n_h, detector pixel (m), spots, depth range (m)
    ("8x8", 100, 8, 8, 6.5e-6, 6, 0.0),
    ("1x300", 100, 1, 300, 6.5e-6, 6, 0.0),
    ("300x3", 100, 300, 3, 6.5e-6, 6, 0.0),
    ("10nm", 1600, 260, 300, 16e-6, 4, 0.0),
    ("2um", 8, 64, 48, 16e-6, 20, 0.0),
    ("500nm", 32, 64, 48, 16e-6, 20, 0.0),
    ("128nm", 125, 130, 70, 16e-6, 10, 0.0),
    ("65nm-3d", 100, 129, 257, 6.5e-6, 30, 1.4e-6),
]


def _scene(engine, configs, n, depth):
    pl = configs.pixel_length
    rng = numpy.random.RandomState(7)
    data = numpy.zeros((n, 5))
    if depth:
        data[:, 0] = rng.uniform(0, depth, n)
    data[:, 1] = rng.uniform(-0.4 * engine.n_w * pl, 0.4 * engine.n_w * pl, n)
    data[:, 2] = rng.uniform(-0.4 * engine.n_h * pl, 0.4 * engine.n_h * pl, n)
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    return data


@pytest.mark.parametrize("precision,tol", [("f64", 1e-11), ("f32", 3e-7)])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_geometry(case, precision, tol):
    _, mag, n_w, n_h, pixel, n, depth = case
    config, configs, params, engine = gpu_engine(YAML % (mag, n_w, n_h, pixel), precision=precision)
    data = _scene(engine, configs, n, depth)
    out = torch.empty((engine.n_w, engine.n_h), dtype=engine.dtype, device=engine.device)
    img, _ = engine.render_expected([(0.033, data)], out=out)
    torch.cuda.synchronize()
    got = img.double().cpu().numpy()
    want, _ = orc.expected_frame([(0.0, data)], params, exposure_time=0.033)
    assert want.max() > 0
    assert abs(got - want).max() <= tol * want.max()
