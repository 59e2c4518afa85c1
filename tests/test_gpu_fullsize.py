"""BASELINE.json's full sizes on the GPU, checked through properties that do not need the
oracle to finish the whole job: order independence, linearity (which lets the oracle check a
full frame through a small subset of its spots), photon conservation, frame-block ==
frame-by-frame, and the detector's moments over four million pixels.

C4: EPI 3-D, 1e5 molecules, 2048 x 2048 CMOS, column FPN.  C5: Gaussian PSF, 1e6 spots,
4096 x 4096."""
import numpy
import pytest
import torch

import epifm_oracle as orc
import scopyon_b200
from conftest import cmos_table, gpu_engine, make_configs
from scopyon_b200.movie import DeviceMovie, frame_block

pytestmark = pytest.mark.gpu

C4 = """
default:
    magnification: 100
    light_source: {angle: {value: 0.0, units: radian}}
    detector: {type: CMOS, image_size: [2048, 2048], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
    effects: {photo_bleaching: {switch: true, half_life: {value: 2.5, units: s}}}
"""
C5 = """
default:
    magnification: 100
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    detector: {type: CMOS, image_size: [4096, 4096], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
"""
PL = 6.5e-8


@pytest.fixture(autouse=True)
def fresh_table_store():
    """Each test starts and ends with an empty PSF-table store (a full 3-D set is 52 GB)."""
    from scopyon_b200.engine import SatStore
    SatStore.clear_shared()
    torch.cuda.empty_cache()
    yield
    SatStore.clear_shared()
    torch.cuda.empty_cache()


def render(engine, data, dtype):
    out = torch.empty((engine.n_w, engine.n_h), dtype=dtype, device=engine.device)
    img, _ = engine.render_expected([(0.033, data)], out=out)
    torch.cuda.synchronize()
    assert int(engine.errors.item()) == 0
    return img.cpu().numpy().astype(numpy.float64)


def c4_scene(n=100000, seed=123):
    rng = numpy.random.RandomState(seed)
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0.0, 1.5e-6, n)              # beyond the 1 um cutoff: the frozen table
    data[:, 1] = rng.uniform(-1024 * PL, 1024 * PL, n)
    data[:, 2] = rng.uniform(-1024 * PL, 1024 * PL, n)
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1.0
    return rng, data


def test_c4_render_full_size_fp32_order_linearity_and_oracle_subset():
    """The production mode of the benchmark: fp32 box tables for all 1003 depth keys (17 GB),
    32-bit strip accumulators."""
    _, configs, params, engine = gpu_engine(C4, precision="f32")
    engine.ensure_all_tables()
    rng, data = c4_scene()
    whole = render(engine, data, torch.float32)
    assert whole.shape == (2048, 2048) and whole.min() >= 0
    peak = whole.max()
    # any order of the spots: the same bits
    assert numpy.array_equal(render(engine, data[rng.permutation(len(data))], torch.float32), whole)
    # two halves add up to the whole (each render picks its own per-strip LSB: fp32 tolerance)
    halves = render(engine, data[:50000], torch.float32) + render(engine, data[50000:], torch.float32)
    assert abs(halves - whole).max() < 1e-6 * peak
    # the oracle checks the full frame through 40 of its spots: frame minus frame-without-them
    pick = rng.choice(len(data), 40, replace=False)
    pick[:4] = numpy.argsort(data[:, 0])[[0, 1, -2, -1]]             # shallowest and deepest molecules
    rest = numpy.ones(len(data), dtype=bool)
    rest[pick] = False
    want, _ = orc.expected_frame([(0.0, data[pick])], params, exposure_time=0.033)
    alone = render(engine, data[pick], torch.float32)
    assert abs(alone - want).max() < 5e-7 * want.max()
    assert ((alone > 0) == (want > 0)).all()                          # identical pixel footprints
    difference = whole - render(engine, data[rest], torch.float32)
    assert abs(difference - want).max() < 1e-6 * peak
    # photon conservation for spots whose footprint is inside the frame: every photon the oracle
    # deposits for a molecule is in the image
    inside = (abs(data[:, 1]) < 1000 * PL) & (abs(data[:, 2]) < 1000 * PL)
    sample = numpy.flatnonzero(inside)[:25]
    want_mass = orc.expected_frame([(0.0, data[sample])], params, exposure_time=0.033)[0].sum()
    assert abs(render(engine, data[sample], torch.float32).sum() - want_mass) < 1e-6 * want_mass


def test_c4_render_full_size_exact_mode_2d():
    """fp64 tables and 64-bit fixed-point accumulators on the full 2048 x 2048 frame with 1e5 spots
    on the focal plane (one table): bitwise order independence, linearity and mass to 1e-11."""
    _, configs, params, engine = gpu_engine(C4, precision="f64")
    rng, data = c4_scene(seed=7)
    data[:, 0] = 0.0
    whole = render(engine, data, torch.float64)
    assert numpy.array_equal(render(engine, data[rng.permutation(len(data))], torch.float64), whole)
    halves = render(engine, data[:30000], torch.float64) + render(engine, data[30000:], torch.float64)
    assert abs(halves - whole).max() < 1e-11 * whole.max()
    inside = (abs(data[:, 1]) < 1000 * PL) & (abs(data[:, 2]) < 1000 * PL)
    only = render(engine, data[inside], torch.float64)
    one, _ = orc.expected_frame([(0.0, data[inside][:1])], params, exposure_time=0.033)
    assert abs(only.sum() - inside.sum() * one.sum()) < 1e-9 * only.sum()      # every molecule: the same weight and table
    pick = rng.choice(len(data), 30, replace=False)
    want, _ = orc.expected_frame([(0.0, data[pick])], params, exposure_time=0.033)
    rest = numpy.ones(len(data), dtype=bool)
    rest[pick] = False
    assert abs(whole - render(engine, data[rest], torch.float64) - want).max() < 1e-9 * whole.max()


def test_c4_detector_full_size_moments_and_determinism():
    """CMOS + column FPN on 2048 x 2048 (the streaming kernel): reproducible per (seed, frame),
    mean and variance of the counts over 4.2e6 pixels as the model predicts."""
    _, configs, params, engine = gpu_engine(C4, precision="f32")
    rng = numpy.random.RandomState(3)
    photons_host = rng.gamma(2.0, 0.4, (2048, 2048)).astype(numpy.float32)       # mean 0.8, a few pixels >> 1
    photons_host[:8] = rng.uniform(20, 400, (8, 2048))                            # bright rows: the general samplers
    photons = torch.from_numpy(photons_host).to(engine.device)
    adc = torch.empty_like(photons)
    engine.detect(photons, 5, 77, adc=adc)
    first = adc.cpu().numpy().astype(numpy.float64)
    engine.detect(photons, 5, 77, adc=adc)
    assert numpy.array_equal(adc.cpu().numpy(), first)
    engine.detect(photons, 6, 77, adc=adc)
    assert not numpy.array_equal(adc.cpu().numpy(), first)
    # model: counts = offset_j + (Poisson(qe (photons + bg)) + read) * (2^16 - offset_j) / fullwell, no clipping here
    values, p = orc.cmos_readout_pmf(cmos_table())
    read_mean, read_var = (values * p).sum(), (values ** 2 * p).sum() - (values * p).sum() ** 2
    offset = engine.offset.cpu().numpy().astype(numpy.float64)[None, :]
    gain = (65536.0 - offset) / 30000.0
    lam = params["QE"] * (photons_host.astype(numpy.float64) + params["background_mean"])
    mean = offset + (lam + read_mean) * gain
    var = (lam + read_var) * gain ** 2
    assert first.min() >= 0 and first.max() <= 65535
    z = (first - mean).sum() / numpy.sqrt(var.sum())
    assert abs(z) < 5
    chi = ((first - mean) ** 2 / var).mean()
    assert abs(chi - 1) < 0.02                                        # heavy-tailed read noise: generous, still 1e-2
    # per-column offsets: each column's mean follows its own offset
    col = (first[8:] - (lam[8:] + read_mean) * gain).mean(axis=0)
    assert abs(col - offset[0]).max() < 6 * numpy.sqrt(var[8:].mean(axis=0).max() / 2040)


def test_c4_movie_full_size_frame_blocks_equal_frame_by_frame():
    """1e5 molecules, 2048 x 2048, bleaching on: a block of 5 frames in one launch per kernel equals
    five single-frame launches, and a rank that replays the prefix renders the same frames."""
    config = scopyon_b200.DefaultConfiguration()
    config.update(C4)
    lower, upper = [-1024 * PL, -1024 * PL, 0.0], [1024 * PL, 1024 * PL, 1.5e-6]

    def movie():
        return DeviceMovie(config, 100000, lower, upper, 1e-13, 123, precision="f32")

    sequential = movie()
    frames = torch.empty((5, 2048, 2048), dtype=torch.float32, device=sequential.engine.device)
    for k in range(5):
        sequential.render_next(frames[k])
    block = torch.empty_like(frames)
    blocked = movie()
    blocked.render_block(block)
    assert torch.equal(block, frames)
    first, last = frame_block(5, 1, 2)
    shard = movie()
    shard.reset(first_frame=first)
    part = torch.empty((last - first, 2048, 2048), dtype=torch.float32, device=shard.engine.device)
    shard.render_block(part)
    assert torch.equal(part, frames[first:last])
    assert 100 < float(frames.mean()) < 115 and not torch.equal(frames[0], frames[1])


def test_c5_gaussian_million_spots_full_size():
    """1e6 Gaussian spots on 4096 x 4096 through the default (box-table) path: order independence and
    photon conservation against the oracle's table integral."""
    _, configs, params, engine = gpu_engine(C5, precision="f32")
    rng = numpy.random.RandomState(11)
    n = 1000000
    data = numpy.zeros((n, 5))
    data[:, 1] = rng.uniform(-2030 * PL, 2030 * PL, n)
    data[:, 2] = rng.uniform(-2030 * PL, 2030 * PL, n)
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1.0
    whole = render(engine, data, torch.float32)
    assert numpy.array_equal(render(engine, data[rng.permutation(n)], torch.float32), whole)
    one, _ = orc.expected_frame([(0.0, data[:1])], params, exposure_time=0.033)
    assert abs(whole.sum() - n * one.sum()) < 2e-6 * whole.sum()      # all footprints are inside the frame
    pick = rng.choice(n, 30, replace=False)
    want, _ = orc.expected_frame([(0.0, data[pick])], params, exposure_time=0.033)
    rest = numpy.ones(n, dtype=bool)
    rest[pick] = False
    assert abs(whole - render(engine, data[rest], torch.float32) - want).max() < 1e-6 * whole.max()
