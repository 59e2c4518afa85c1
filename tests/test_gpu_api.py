"""The drop-in public API on the GPU: form_image / generate_images / output_frame with the
reference's signatures, checked against the golden vectors and the oracle."""
import warnings

import numpy
import pytest

import epifm_oracle as orc
import scopyon_b200
from conftest import format_inputs, golden, make_configs

pytestmark = pytest.mark.gpu


def tirf_config():
    config = scopyon_b200.DefaultConfiguration()
    config.default.detector.exposure_time = 33.0e-3
    return config


def test_form_image_tirf_example(known_answers):
    """examples/tirf.py / detection.py verbatim (BASELINE config 1)."""
    g = golden("tirf_c1.npz")
    config = tirf_config()
    rng = numpy.random.RandomState(123)
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    inputs = rng.uniform(-L_2, +L_2, size=(100, 2))
    assert numpy.array_equal(inputs, g["inputs"])
    img, info = scopyon_b200.form_image(inputs, config=config, rng=rng, full_output=True)
    assert isinstance(img, scopyon_b200.Image) and img.shape == (512, 512) and img.dtype == numpy.float64
    want = 0.92 * (g["photons"] + 0.01)
    got = info["expectation"]
    assert abs(got - want).max() / want.max() < 1e-6           # fp32 frame storage
    assert abs(got.sum() - known_answers["tirf_c1"]["expectation_sum"]) / got.sum() < 1e-6
    # true_data: [t, state, X px, Y px, X m, Y m, depth*t, normalization]  
    t0 = info["true_data"][0]
    assert numpy.allclose(t0, [0.033, 1, 356.092223, 146.003339, 6.67832186e-06, -7.26948783e-06, 0, 112.385481],
                          rtol=1e-8)
    assert len(info["true_data"]) == 100 and "fluorescence_states" not in info
    # EMCCD frame statistics: offset 2000 counts, gain 12.59 e-/count, readout 100 e-, EM gain 300
    adc = img.as_array()
    dark = want < 0.0093
    assert abs(adc[dark].mean() - (2000 + 0.0092 * 300 / 12.591286829513976)) < 0.2
    assert abs(adc[dark].std() - numpy.sqrt(100 ** 2 + 2 * 0.0092 * 300 ** 2) / 12.591286829513976) < 0.3
    bright = want > 4
    assert adc[bright].mean() > 2000 + 0.8 * want[bright].mean() * 300 / 12.59
    # plain call returns just the image; a seeded rng reproduces it
    again = scopyon_b200.form_image(inputs, config=config, rng=restart(123, 200))
    assert numpy.array_equal(again.as_array(), adc)
    other = scopyon_b200.form_image(inputs, config=config, rng=numpy.random.RandomState(5))
    assert not numpy.array_equal(other.as_array(), adc)


def restart(seed, n_uniform):
    rng = numpy.random.RandomState(seed)
    rng.uniform(size=n_uniform)
    return rng


def test_twocolor_example():
    """examples/twocolor.py (BASELINE config 2): two channels from one rng, stacked as RGB."""
    config = scopyon_b200.DefaultConfiguration()
    config.update("default: {detector: {type: CCD, image_size: [128, 128]}}")
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = 128 * pixel_length * 0.5
    rng = numpy.random.RandomState(123)
    inputs = rng.uniform(-L_2, +L_2, size=(250, 2))
    img1 = scopyon_b200.form_image(inputs[: 200], config=config, rng=rng)
    img2 = scopyon_b200.form_image(inputs[150:], config=config, rng=rng)
    img = scopyon_b200.Image.RGB(red=img1, green=img2)
    assert img.shape == (128, 128, 3) and (img.as_array()[:, :, 2] == 0).all()
    assert numpy.array_equal(img.as_array()[:, :, 0], img1.as_array())
    assert img.as_8bit().dtype == numpy.uint8


def test_output_frame_with_state_dict_against_oracle():
    """Motion blur + photobleaching through the reference's seam (_epifm.py:1121-1225):
    budgets handed in as a dict, updated in place, expectation / true_data / states equal
    the oracle's frame by frame."""
    g = golden("movie_ccd.npz")
    config, configs, params = make_configs(str(g["yaml"]))
    inputs = [(float(t), p) for t, p in zip(g["times"], g["points"])]
    data = format_inputs(config, inputs)
    from scopyon_b200._epifm import _EPIFMSimulator
    from scopyon_b200.engine import DeviceEngine
    sim = _EPIFMSimulator(configs)
    sim._engine = DeviceEngine(configs, precision="f64")
    rng = numpy.random.RandomState(0)
    ids = g["true_ids"]
    budgets = {int(i): float(b) for i, b in zip(ids, numpy.linspace(0.05, 2.5, len(ids)))}   # some bleach mid-movie
    mine, theirs = dict(budgets), dict(budgets)
    psf = orc.PsfTables(params)
    for f in range(4):
        camera, info = sim.output_frame(data, frame_index=f, fluorescence_states=mine, rng=rng)
        photons, true_data = orc.expected_frame(data, params, frame_index=f, fluorescence_states=theirs, psf=psf)
        want = orc.detector_expectation(photons, params)
        assert abs(camera[:, :, 0] - want).max() / want.max() < 1e-9
        assert set(info["true_data"]) == set(true_data)
        for m in true_data:
            assert numpy.allclose(info["true_data"][m], true_data[m], rtol=1e-12, atol=0)
        assert set(mine) == set(theirs)
        assert all(abs(mine[m] - theirs[m]) <= 1e-12 * max(theirs[m], 1e-30) for m in theirs)
        assert info["fluorescence_states"] == mine
    assert sum(1 for b in mine.values() if b == 0) >= 2        # some molecules did bleach


def test_generate_images_movie_and_bleaching_decay():
    """examples/bleaching.py shape (BASELINE config 3), shortened: N=300, 12 frames."""
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 360
    detector: {exposure_time: 0.033, image_size: [256, 256], type: CCD}
    effects: {photo_bleaching: {switch: true, half_life: 0.1}}
""")
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = 256 * pixel_length * 0.5
    rng = numpy.random.RandomState(123)
    num_frames, dt = 12, 0.033
    t = numpy.arange(0, (num_frames + 1) * dt, dt)
    inputs = scopyon_b200.sample_inputs(t, N=300, lower=-L_2, upper=+L_2, ndim=2, D=0.1e-12, rng=rng)
    gen = scopyon_b200.generate_images(inputs, num_frames=num_frames, config=config, rng=rng, full_output=True)
    assert hasattr(gen, "__next__")                            # a generator like the reference's
    frames = list(gen)
    assert len(frames) == num_frames
    totals = [info["expectation"].sum() - 0.92 * 0.01 * 256 * 256 for _, info in frames]
    alive = [sum(1 for b in info["fluorescence_states"].values() if b > 0) for _, info in frames]
    assert all(a >= b for a, b in zip(alive, alive[1:])) and alive[-1] < 0.35 * 300 < alive[0]
    assert totals[-1] < 0.4 * totals[0]
    # survival follows the half-life: after k frames exp(-ln2 * k dt / T) of the molecules emit
    want = 300 * numpy.exp(-numpy.log(2) * num_frames * dt / 0.1)
    assert abs(alive[-1] - want) < 5 * numpy.sqrt(want) + 5
    imgs = list(scopyon_b200.generate_images(inputs, num_frames=3, config=config, rng=numpy.random.RandomState(1)))
    assert all(isinstance(i, scopyon_b200.Image) for i in imgs)


def test_cmos_3d_epi_movie_against_oracle_expectation():
    """Scaled-down BASELINE config 4: EPI 3-D, CMOS, column FPN; bleaching off so the
    expectation is deterministic and comparable to the golden reference frames."""
    g = golden("movie_cmos3d.npz")
    yaml = str(g["yaml"]) + "    effects: {photo_bleaching: {switch: false}}\n"
    config = scopyon_b200.DefaultConfiguration()
    config.update(yaml)
    inputs = [(float(t), p) for t, p in zip(g["times"], g["points"])]
    frames = list(scopyon_b200.generate_images(inputs, num_frames=2, config=config, rng=numpy.random.RandomState(3),
                                               full_output=True))
    _, _, params = make_configs(yaml)
    data = format_inputs(config, inputs)
    for f, (img, info) in enumerate(frames):
        photons, _ = orc.expected_frame(data, params, frame_index=f)
        want = orc.detector_expectation(photons, params)
        assert abs(info["expectation"] - want).max() / want.max() < 1e-6
        assert "fluorescence_states" not in info
        adc = img.as_array()
        assert adc.min() >= 0 and abs(adc.mean() - (100 + (want.mean() + 1.9) / (30000 / (65536 - 100)))) < 3


@pytest.mark.parametrize("full_output", [False, True])
def test_large_frames_leave_the_device_as_float32_and_equal_the_device_widened_ones(full_output, monkeypatch):
    """Frames of >= 1024 x 1024 pixels leave the device as float32: into the page-locked payload of
    the Image (float64 array made on demand, image.HostPlane) or -- when the caller already holds
    LAZY_PINNED_BYTES of payloads -- into staging memory widened by host threads (scb_host_widen_*),
    three frames in flight; smaller ones are widened on the device.  Same seeds, same kernels: the
    float64 arrays must be identical, frame for frame, on all three routes."""
    from scopyon_b200 import engine as engine_module

    def movie(min_pixels, lazy_bytes):
        monkeypatch.setattr(engine_module, "HOST_WIDEN_MIN_PIXELS", min_pixels)
        monkeypatch.setattr(engine_module, "LAZY_PINNED_BYTES", lazy_bytes)
        config = scopyon_b200.DefaultConfiguration()
        config.update("""
default:
    magnification: 100
    detector: {type: CMOS, image_size: [1024, 1040], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
""")
        rng = numpy.random.RandomState(5)
        pl = 6.5e-6 / 100
        points = numpy.stack([rng.uniform(-500 * pl, 500 * pl, 400), rng.uniform(-500 * pl, 500 * pl, 400)], axis=1)
        inputs = [(k * 0.033, points + k * 1e-8) for k in range(8)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(9))
            out = list(sim.generate_images(inputs, num_frames=7, full_output=full_output))
        if full_output:
            return [img for img, _ in out], [info["expectation"].copy() for _, info in out]
        return out, []

    lazy_images, lazy_expect = movie(1 << 20, 1 << 30)
    if not full_output:
        # the Image carries the float32 frame; the float64 array appears on demand and is exact
        payload = lazy_images[0].as_array(numpy.float32)
        assert payload.dtype == numpy.float32 and lazy_images[0].dtype == numpy.float64
        assert numpy.array_equal(lazy_images[0].as_array(), payload.astype(numpy.float64))
    lazy_frames = [img.as_array().copy() for img in lazy_images]
    host_images, host_expect = movie(1 << 20, 0)
    host_frames = [img.as_array().copy() for img in host_images]
    device_images, device_expect = movie(1 << 40, 1 << 30)
    device_frames = [img.as_array().copy() for img in device_images]
    assert len(host_frames) == 7 and host_frames[0].dtype == numpy.float64 and host_frames[0].shape == (1024, 1040)
    for a, b, c in zip(host_frames + host_expect, device_frames + device_expect, lazy_frames + lazy_expect):
        assert numpy.array_equal(a, b) and numpy.array_equal(a, c)
    assert not numpy.array_equal(host_frames[0], host_frames[1])
    assert 100 < host_frames[0].mean() < 120


def test_generator_abandoned_with_frames_in_flight():
    """Breaking out of generate_images leaves two frames enqueued: closing the generator waits for
    their downloads and widening jobs, and the next movie is unaffected (same seeds, same frames)."""
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 100
    detector: {type: CMOS, image_size: [1024, 1024], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
""")
    rng = numpy.random.RandomState(2)
    points = rng.uniform(-400 * 6.5e-8, 400 * 6.5e-8, (300, 2))
    inputs = [(k * 0.033, points + k * 2e-8) for k in range(10)]

    def first_frames(count, stop_after):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(4))
            frames = []
            gen = sim.generate_images(inputs, num_frames=count)
            for img in gen:
                frames.append(img.as_array().copy())
                if len(frames) == stop_after:
                    break
            gen.close()
        return frames

    short = first_frames(9, 2)            # frames 2 and 3 are in flight when the loop stops
    full = first_frames(9, 9)
    assert len(short) == 2 and len(full) == 9
    assert numpy.array_equal(short[0], full[0]) and numpy.array_equal(short[1], full[1])


def test_device_resident_trajectories_give_the_same_movie():
    """sample_inputs(..., device=True) leaves the trajectory on the GPU (DevicePoints); generate_images and
    form_image render it in place -- the same numbers as the host arrays, hence the same frames, budgets
    and true_data, bit for bit."""
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 100
    detector: {type: CMOS, image_size: [128, 96], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    effects: {photo_bleaching: {switch: true, half_life: {value: 0.1, units: s}}}
""")
    pl = 6.5e-8
    t = numpy.arange(0, 6) * 0.033
    kwargs = dict(N=300, lower=[-60 * pl, -45 * pl, 0.0], upper=[60 * pl, 45 * pl, 3e-7], D=2e-13, ndim=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        host = scopyon_b200.sample_inputs(t, rng=numpy.random.RandomState(21), **kwargs)
        resident = scopyon_b200.sample_inputs(t, rng=numpy.random.RandomState(21), device=True, **kwargs)
        assert all(numpy.array_equal(a[1], b[1].cpu()) and a[0] == b[0] for a, b in zip(host, resident))
        movies = []
        for inputs in (host, resident):
            sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(4))
            movies.append(list(sim.generate_images(inputs, num_frames=5, full_output=True)))
        single = [scopyon_b200.form_image(inputs[2][1], config=config, rng=numpy.random.RandomState(6)).as_array()
                  for inputs in (host, resident)]
    for (img_a, info_a), (img_b, info_b) in zip(*movies):
        assert numpy.array_equal(img_a.as_array(), img_b.as_array())
        assert numpy.array_equal(info_a["expectation"], info_b["expectation"])
        assert info_a["fluorescence_states"] == info_b["fluorescence_states"]
        assert info_a["true_data"].keys() == info_b["true_data"].keys()
        assert all(numpy.array_equal(info_a["true_data"][k], info_b["true_data"][k]) for k in info_a["true_data"])
    assert numpy.array_equal(single[0], single[1]) and single[0].max() > 110


@pytest.mark.parametrize("device_inputs", [False, True])
def test_blocks_of_frames_equal_the_frame_by_frame_movie(device_inputs, monkeypatch):
    """generate_images hands the engine blocks of BLOCK_FRAMES single-snapshot frames (one binning / render /
    detector launch per block, engine.begin_block); the frames are those of the frame-by-frame route, bit for
    bit -- TIRF illumination and bleaching, so weights differ from molecule to molecule and frame to frame, a
    last block that is not full, and a frame of two snapshots (motion blur) in the middle that must go alone."""
    from scopyon_b200 import engine as engine_module
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 100
    detector: {type: CMOS, image_size: [1024, 1024], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
    effects: {photo_bleaching: {switch: true, half_life: {value: 0.3, units: s}}}
""")
    pl = 6.5e-8
    t = list(numpy.arange(0, 23) * 0.033)
    t.insert(12, 11.5 * 0.033)          # frame 11 sees two snapshots
    kwargs = dict(N=600, lower=[-480 * pl, -480 * pl, 0.0], upper=[480 * pl, 480 * pl, 4e-7], D=3e-13, ndim=3)

    def movie(block_frames):
        monkeypatch.setattr(engine_module, "BLOCK_FRAMES", block_frames)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            inputs = scopyon_b200.sample_inputs(numpy.array(t), rng=numpy.random.RandomState(3), device=device_inputs,
                                                **kwargs)
            sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(8))
            return [img.as_array(numpy.float32).copy() for img in sim.generate_images(inputs, num_frames=21)]

    blocks, singles = movie(8), movie(1)
    assert len(blocks) == 21 and blocks[0].shape == (1024, 1024)
    for k, (a, b) in enumerate(zip(blocks, singles)):
        assert numpy.array_equal(a, b), k
    assert not numpy.array_equal(blocks[0], blocks[1]) and blocks[0].max() > 110


def test_blocks_of_frames_with_ids_that_change_from_frame_to_frame(monkeypatch):
    """The molecules of a movie may be listed in a different order in every snapshot (rows carry their ids):
    each frame of a block then needs its own row -> budget-slot map, and photobleaching must follow the
    molecule, not the row.  Block route == frame-by-frame route, bit for bit."""
    from scopyon_b200 import engine as engine_module
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 100
    detector: {type: CMOS, image_size: [1024, 1024], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    effects: {photo_bleaching: {switch: true, half_life: {value: 0.2, units: s}}}
""")
    pl = 6.5e-8
    rng = numpy.random.RandomState(17)
    n = 500
    base = numpy.zeros((n, 5))
    base[:, 0] = rng.uniform(-450 * pl, 450 * pl, n)
    base[:, 1] = rng.uniform(-450 * pl, 450 * pl, n)
    base[:, 3] = numpy.arange(n)
    base[:, 4] = 1.0
    inputs = []
    for k in range(12):
        rows = base.copy()
        rows[:, :2] += rng.normal(0, 2e-8, (n, 2))
        inputs.append((k * 0.033, rows[rng.permutation(n)]))

    def movie(block_frames):
        monkeypatch.setattr(engine_module, "BLOCK_FRAMES", block_frames)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = scopyon_b200.create_simulator(config, rng=numpy.random.RandomState(8))
            return [img.as_array(numpy.float32).copy() for img in sim.generate_images(inputs, num_frames=11)]

    blocks, singles = movie(8), movie(1)
    for k, (a, b) in enumerate(zip(blocks, singles)):
        assert numpy.array_equal(a, b), k
    assert blocks[10].mean() < blocks[0].mean()          # the molecules bleach
