"""Host-side logic of the drop-in boundary (no GPU)."""
import warnings

import numpy
import pytest

import epifm_oracle as orc
import scopyon_b200
from conftest import format_inputs, golden, make_configs
from scopyon_b200 import _epifm
from scopyon_b200.engine import walker_alias


def test_format_data_matches_reference():
    g = golden("format_data.npz")
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
preprocessing:
    scale: {value: 1.0e-6, units: m}
    origin: {value: [1.0e-6, -2.0e-6, 0.5e-6], units: m}
    unit_x: {value: [0.0, 1.0, 0.0], units: m}
    unit_y: {value: [0.0, 0.0, 1.0], units: m}
""")
    for width in (2, 3, 4, 5):
        got = format_inputs(config, g["in{}".format(width)])[0][1]
        assert numpy.array_equal(got, g["out{}".format(width)])


def test_input_errors_like_reference():
    config = scopyon_b200.DefaultConfiguration()
    with pytest.raises(ValueError):
        format_inputs(config, numpy.zeros((3, 6)))
    with pytest.raises(ValueError):
        format_inputs(config, numpy.zeros(3))
    with pytest.raises(ValueError):
        format_inputs(config, [(0.0, [1, 2, 3])])
    with pytest.raises(TypeError):
        format_inputs(config, 3.0)
    with pytest.raises(TypeError):
        scopyon_b200.EPIFMSimulator(config=3, rng=numpy.random.RandomState(0))
    with pytest.warns(UserWarning):
        scopyon_b200.EPIFMSimulator(config=config)


def test_configs_flatten_known_answers(known_answers):
    _, configs, params = make_configs("default: {detector: {exposure_time: 0.033}}")
    want = known_answers["scalars"]
    amplitude, depth = configs.snells_law()
    assert amplitude == want["snells_amplitude"] and depth == want["snells_depth"]
    phys = configs.photophysics()
    n_emit = phys.quantum_yield * (phys.amplitude0 * phys.x_sec * 0.033) * phys.absorb_frac
    assert n_emit == want["n_emit_33ms"]
    beta, n_emit0 = orc.photon_budget_scale(params)
    assert phys.budget_scale == beta * n_emit0
    assert configs.n_radial() == 1000 and configs.n_depth_keys() == 1002
    geom = configs.geometry()
    assert geom.pixel_length == 16e-6 / 241.0 and geom.n_w == 512
    # epi-illumination below the critical angle
    _, epi, _ = make_configs("default: {light_source: {angle: {value: 0.0, units: radian}}}")
    assert epi.snells_law()[1] == numpy.inf


def test_gaussian_and_unsupported_settings():
    _, configs, _ = make_configs("""
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
""")
    assert configs.psf_radial_width == 100e-9 and configs.fluoem_norm_sum == 1.0
    assert abs(configs.psf_wavelength - 600e-9) < 1e-12
    with pytest.raises(NotImplementedError):
        make_configs("default: {dichroic_mirror: {switch: true}}")
    with pytest.raises(ValueError):
        make_configs("default: {analog_to_digital_converter: {type: row}}")
    with pytest.raises(ValueError):
        make_configs("default: {fluorophore: {type: NoSuchDye}}")
    with pytest.raises(ValueError):
        make_configs("default: {type: confocal}")


def test_frame_windows_match_oracle():
    _, configs, params = make_configs("default: {detector: {exposure_time: 0.033}}")
    times = numpy.arange(0, 0.2, 0.011)
    for frame in range(5):
        for start in (0.0, 0.004, 0.05):
            got, t, exposure = _epifm.frame_windows(times, frame, start, 0.033, configs)
            want, exposure_o = orc.frame_windows(times, frame, start, 0.033, params)
            assert got == [(k, float(u)) for k, u in want] and exposure == exposure_o
    # a single snapshot at t = 0 covers the whole exposure (form_image)
    got, _, _ = _epifm.frame_windows(numpy.array([0.0]), 0, 0.0, 0.1, configs)
    assert got == [(0, 0.1)]


def test_depth_keys_vectorised_rule():
    depths = numpy.array([0.0, 0.29e-6, -5.5e-9, 1.0009e-6, 1.0011e-6, 3e-6, 0.9999999e-6])
    keys = _epifm.depth_keys_of(depths, 1000e-9, 1002)
    want = [orc.depth_key(d, 1000e-9)[0] for d in depths]
    assert [int(k) if k < 1002 else -1 for k in keys] == want


def test_walker_alias_reproduces_distribution():
    rn = _epifm.catalog_tables()["cmos_readout"]
    table = walker_alias(rn["electrons"], rn["weight"]).astype(numpy.float64)
    n = len(table)
    p = numpy.asarray(rn["weight"]) / numpy.sum(rn["weight"])
    values = numpy.asarray(rn["electrons"])
    mass = {}
    for value, alias_value, threshold, _ in table:
        mass[value] = mass.get(value, 0.0) + threshold / n
        mass[alias_value] = mass.get(alias_value, 0.0) + (1.0 - threshold) / n
    got = numpy.array([mass.get(numpy.float32(v).astype(numpy.float64), 0.0) for v in values])
    assert abs(got - p).max() < 1e-7
    assert n == 194 and abs(got.sum() - 1.0) < 1e-6


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    config = scopyon_b200.DefaultConfiguration()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            scopyon_b200.form_image(numpy.zeros((3, 2)), config=config, rng=numpy.random.RandomState(0))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            scopyon_b200.sample_inputs([0.0, 0.1], N=4, ndim=2, rng=numpy.random.RandomState(0))


def test_product_never_imports_oracle():
    import os
    from conftest import ROOT
    pkg = os.path.join(ROOT, "scopyon_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "epifm_oracle" not in text and "c_oracle" not in text and "ref_shim" not in text, name


def test_image_with_float32_payload_is_the_float64_image():
    """Frames that came off the GPU as float32 (image.HostPlane) behave as the float64 arrays the
    reference hands out: dtype, shape, as_array, 8-bit scaling, save; the payload is widened once."""
    from scopyon_b200.image import HostPlane, Image
    rng = numpy.random.RandomState(2)
    f32 = (rng.standard_normal((12, 17)) * 100 + 100).astype(numpy.float32)
    calls = []

    def widen(src):
        calls.append(1)
        return src.astype(numpy.float64)

    img = Image(HostPlane(f32, widen=widen))
    plain = Image(f32.astype(numpy.float64))
    assert img.dtype == numpy.float64 and img.shape == (12, 17) and img.ndim == 2 and img.size == 12 * 17
    assert img.as_array(numpy.float32) is f32 and not calls          # no widening for the float32 payload
    a = img.as_array()
    assert a.dtype == numpy.float64 and numpy.array_equal(a, plain.as_array()) and len(calls) == 1
    assert img.as_array() is a and len(calls) == 1                    # kept
    assert numpy.array_equal(img.as_8bit().as_array(), plain.as_8bit().as_array())
    rgb = Image.RGB(red=Image(HostPlane(f32)), green=plain)
    assert rgb.shape == (12, 17, 3) and numpy.array_equal(rgb.as_array()[:, :, 0], rgb.as_array()[:, :, 1])
    assert numpy.array_equal(numpy.asarray(HostPlane(f32)), plain.as_array())


def test_streaming_npy_writer(tmp_path):
    """NpyFrameWriter: blocks appended in order give the file numpy.save writes for the whole stack."""
    from scopyon_b200.movie import NpyFrameWriter
    rng = numpy.random.RandomState(4)
    for dtype in (numpy.float32, numpy.uint16, numpy.uint8):
        stack = (rng.uniform(0, 200, (11, 6, 9))).astype(dtype)
        path = tmp_path / "movie_{}.npy".format(numpy.dtype(dtype).name)
        with NpyFrameWriter(path, 11, (6, 9), dtype) as writer:
            writer(0, stack[:4])
            writer(4, stack[4:8])
            writer(8, stack[8:])
        assert numpy.array_equal(numpy.load(path), stack) and numpy.load(path).dtype == dtype
    with pytest.raises(ValueError):
        with NpyFrameWriter(tmp_path / "short.npy", 5, (6, 9), numpy.float32) as writer:
            writer(0, numpy.zeros((3, 6, 9), numpy.float32))                 # 3 of 5 frames
    with pytest.raises(ValueError):
        with NpyFrameWriter(tmp_path / "wrong.npy", 5, (6, 9), numpy.float32) as writer:
            writer(0, numpy.zeros((3, 6, 9), numpy.float64))                 # wrong dtype
