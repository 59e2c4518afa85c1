"""Tensor-core renderer for the separable Gaussian PSF (C ABI: scb_render_gaussian_tc).

Tolerance: 1e-5 of the image maximum (north_star) at sigma = 100 nm.  The reference table
(radial profile, linearly interpolated, clamped in the corners) is not exactly separable; the
product of its own marginals reproduces it to 1.1e-6 of the peak at sigma = 100 nm and to
1.1e-5 at sigma = 200 nm (corner clamp), the split-tf32 contraction adds ~1e-6.  The exact
SAT path is the yardstick."""
import numpy
import pytest
import torch

from conftest import format_inputs, golden, make_configs

pytestmark = pytest.mark.gpu

GAUSS = """
default:
    fluorophore: {type: Gaussian, radial_width: {value: %g, units: m}, wave_length: {value: 600.0e-9, units: m}}
    magnification: 100
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: 6.5e-6, units: m}, exposure_time: 0.033}
"""


def engines(yaml):
    from scopyon_b200.engine import DeviceEngine, SatStore
    config, configs, params = make_configs(yaml)
    SatStore.clear_shared()
    return config, configs, DeviceEngine(configs, precision="f64", gaussian_tc=False), \
        DeviceEngine(configs, precision="f64", gaussian_tc=True)


def render(engine, data, dtype=torch.float64):
    out = torch.empty((engine.n_w, engine.n_h), dtype=dtype, device=engine.device)
    img, _ = engine.render_expected([(0.033, data)], out=out)
    torch.cuda.synchronize()
    assert int(engine.errors.item()) == 0
    return img.cpu().numpy()


def scene(configs, n, size, rng, spread=1.05):
    pl = configs.pixel_length
    data = numpy.zeros((n, 5))
    data[:, 1] = rng.uniform(-size[0] * pl / 2 * spread, size[0] * pl / 2 * spread, n)
    data[:, 2] = rng.uniform(-size[1] * pl / 2 * spread, size[1] * pl / 2 * spread, n)
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    return data


def test_matches_reference_golden_case():
    g = golden("gaussian_case.npz")
    yaml = """
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    detector: {image_size: [64, 64], exposure_time: 0.033}
"""
    config, configs, exact, tc = engines(yaml)
    assert tc.gaussian_tc and not exact.gaussian_tc
    data = format_inputs(config, g["inputs"])[0][1]
    got = render(tc, data)
    want = g["photons"]                                   # the live reference's image
    assert abs(got - want).max() / want.max() < 1e-5
    assert abs(got.sum() - want.sum()) / want.sum() < 1e-5


@pytest.mark.parametrize("sigma,tol", [(100e-9, 5e-6), (200e-9, 2e-5)])
def test_random_scene_against_exact_sat_path(sigma, tol):
    """4000 spots on 500 x 390 (ragged 128-pixel tiles), spots hanging over the border."""
    config, configs, exact, tc = engines(GAUSS % (sigma, 500, 390))
    rng = numpy.random.RandomState(3)
    data = scene(configs, 4000, (500, 390), rng)
    data[::7, 4] = 0                                      # dark molecules
    want = render(exact, data)
    got = render(tc, data)
    assert abs(got - want).max() / want.max() < tol
    assert abs(got.sum() - want.sum()) / want.sum() < 2e-6
    got32 = render(tc, data, dtype=torch.float32)
    assert abs(got32 - want).max() / want.max() < tol
    assert numpy.array_equal(render(tc, data), got)        # reproducible run to run
    assert (render(tc, data[:0]) == 0).all()               # empty scene


def test_dense_tile_many_chunks():
    """6000 spots inside one 128 x 128 tile: > 4096 per tile (two sorted segments, 375 chunks)."""
    config, configs, exact, tc = engines(GAUSS % (100e-9, 256, 256))
    rng = numpy.random.RandomState(4)
    pl = configs.pixel_length
    data = scene(configs, 6000, (256, 256), rng)
    data[:, 1] = rng.uniform(-120 * pl, -10 * pl, 6000)
    data[:, 2] = rng.uniform(10 * pl, 120 * pl, 6000)
    want = render(exact, data)
    got = render(tc, data)
    # 6000 products per pixel accumulate in fp32 inside the tensor core: ~sqrt(K) * 2^-24
    assert abs(got - want).max() / want.max() < 1e-5


def test_not_used_for_born_wolf():
    from scopyon_b200.engine import DeviceEngine
    _, configs, _ = make_configs()
    assert DeviceEngine(configs, gaussian_tc=True).gaussian_tc is False
