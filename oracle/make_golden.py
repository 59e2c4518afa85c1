"""Generate ``tests/golden/*`` from the LIVE, unmodified reference (build container only).

The reference holds no golden image vectors; its own tests only
print six PSF integrals.  This script runs the reference itself (through
``oracle/ref_shim.py``) on small seeded cases and stores inputs + outputs, so that the
oracle and the CUDA path can be checked on machines where ``/root/reference`` does not
exist (the GPU box).  Run:  ``python oracle/make_golden.py``
"""
import json
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def psf_known_answers(ref):
    """The integrals printed by the reference's test/test_epifm.py:16-70, plus table samples."""
    from scopyon._epifm import PointSpreadingFunction
    out = {}
    radial = numpy.arange(0.0, 1000.0e-9, 1.0e-9, dtype=float)
    for name, kwargs in (
            ("tritc", dict(psf_radial_width=None, fluorophore_type="Tetramethylrhodamine(TRITC)", psf_wavelength=5.78e-07)),
            ("gaussian", dict(psf_radial_width=1.0e-7, fluorophore_type="Gaussian", psf_wavelength=6.0e-7))):
        psf = PointSpreadingFunction(psf_radial_cutoff=1000.0e-9, psf_depth_cutoff=1000.0e-9, **kwargs)
        psf_r = psf.get_distribution(radial, 0.0)
        cart = psf.radial_to_cartesian(radial, psf_r, 1000.0e-9, 1.0e-9)
        camera = numpy.zeros((512, 512))
        psf.overlay_signal_(camera, cart, numpy.zeros(3, dtype=float), 4.444444444444444e-08, 1.0e-9, 1.0)
        out[name] = dict(
            radial_integral=float(numpy.sum(2 * numpy.pi * radial * psf_r) * 1.0e-9),
            cartesian_integral=float(cart.sum() * 1.0e-18),
            overlay_sum=float(camera.sum()),
            footprint_pixels=int((camera > 0).sum()),
            table_centre=float(cart[999, 999]), table_corner=float(cart[0, 0]))
    return out


def radial_profiles(ref, config):
    from scopyon import _epifm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ec = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    radial = numpy.arange(0.0, 1000.0e-9, 1.0e-9, dtype=float)
    depths = numpy.array([0.0, 1e-9, 100e-9, 289e-9, 500e-9, 800e-9, 1000e-9, 1000.0e-9 + 0.0])
    prof = numpy.stack([ec.fluorophore_psf.get_distribution(radial, z) for z in depths])
    integrals = []
    for z in (0.0, 100e-9, 500e-9, 1000e-9):
        integrals.append(float(ec.fluorophore_psf.get(z).sum() * 1e-18))
    return dict(depths=depths, born_wolf=prof, psf_wavelength=ec.psf_wavelength,
                table_integrals=numpy.array(integrals))


def scalar_known_answers(ref, config):
    from scopyon import _epifm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ec = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    sim = _epifm._EPIFMSimulator(ec, environ=None)
    amplitude, depth = sim.snells_law()
    n_emit = _epifm._EPIFMSimulator.get_emit_photons(amplitude, 0.033, 83400, 0.61, 20e-9)
    return dict(snells_amplitude=float(amplitude), snells_depth=float(depth), n_emit_33ms=float(n_emit),
                psf_wavelength=float(ec.psf_wavelength), fluoem_norm_sum=float(ec.fluoem_norm.sum()),
                adc_gain_none=float(ec.ADConverter_gain[0, 0]), hc=float(ref.constants.hc),
                N_A=float(ref.constants.N_A))


def expectation_case(ref, config, inputs, unit_time):
    """Pre-noise photon image straight from the reference's get_molecule_plane."""
    from scopyon import _epifm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = ref.EPIFMSimulator(config=config, method="default", rng=numpy.random.RandomState(0))
        data = sim._EPIFMSimulator__format_inputs(inputs)
        base = sim.base()
    p_0 = numpy.asarray(base.configs.detector_focal_point)
    shape = tuple(base.configs.detector_image_size)
    true_data = {}
    expected, optinfo, _ = base.get_molecule_plane(
        data[0][1], shape=shape, p_b=p_0, p_0=p_0, unit_time=unit_time,
        optional_info=true_data, fluorescence_states=None, rng=None, processes=1)
    return expected, data[0][1]


def movie_case(ref, yaml_update, n, frames, ndim, seed, size, sub_steps=3):
    config = ref.DefaultConfiguration()
    config.update(yaml_update)
    pl = config.default.detector.pixel_length / config.default.magnification
    L = size[0] * pl * 0.5
    rng0 = numpy.random.RandomState(seed)
    dt = config.default.detector.exposure_time
    t = numpy.arange(0, (frames + 1) * dt, dt / sub_steps)
    if ndim == 2:
        inputs = ref.sample_inputs(t, N=n, lower=-L, upper=L, ndim=2, D=1e-13, rng=rng0)
    else:
        inputs = ref.sample_inputs(t, N=n, lower=[-L, -L, 0], upper=[L, L, 1.4e-6], ndim=3, D=1e-13, rng=rng0)
    rng = numpy.random.RandomState(seed + 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = list(ref.generate_images(inputs, num_frames=frames, config=config, rng=rng, full_output=True))
    ids = sorted(out[0][1]["true_data"].keys())
    case = dict(
        yaml=numpy.array(yaml_update), seed=seed, times=numpy.array([tt for tt, _ in inputs]),
        points=numpy.stack([p for _, p in inputs]),
        adc=numpy.stack([img.as_array() for img, _ in out]),
        expectation=numpy.stack([info["expectation"] for _, info in out]),
        true_ids=numpy.array(ids),
        true_data=numpy.stack([[info["true_data"][i] for i in ids] for _, info in out]),
    )
    if "fluorescence_states" in out[0][1]:
        case["budgets"] = numpy.stack([[info["fluorescence_states"][i] for i in ids] for _, info in out])
    return case


def emccd_pmfs(ref):
    from scopyon._epifm import EMCCD
    out = {}
    for i, E in enumerate([0.0092, 0.5, 5.0, 50.0]):
        S, p = EMCCD.probability_distribution(E, 300)
        # store the cdf on a thinned grid (the pmf itself has up to 3e4 entries)
        cdf = numpy.cumsum(p)
        step = max(1, len(S) // 2000)
        out["E{}".format(i)] = numpy.array(E)
        out["S{}".format(i)] = S[::step]
        out["cdf{}".format(i)] = cdf[::step]
        out["p0_{}".format(i)] = numpy.array(p[0] if S[0] == 0 else 0.0)
        out["mean{}".format(i)] = numpy.array((S * p).sum())
        out["var{}".format(i)] = numpy.array((S * S * p).sum() - (S * p).sum() ** 2)
        out["support{}".format(i)] = numpy.array([S[0], S[-1]])
    return out


def format_cases(ref):
    config = ref.DefaultConfiguration()
    config.update("""
preprocessing:
    scale: {value: 1.0e-6, units: m}
    origin: {value: [1.0e-6, -2.0e-6, 0.5e-6], units: m}
    unit_x: {value: [0.0, 1.0, 0.0], units: m}
    unit_y: {value: [0.0, 0.0, 1.0], units: m}
""")
    rng = numpy.random.RandomState(3)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = ref.EPIFMSimulator(config=config, method="default", rng=numpy.random.RandomState(0))
        for width in (2, 3, 4, 5):
            pts = rng.uniform(-3, 3, size=(7, width))
            if width >= 4:
                pts[:, -2] = numpy.arange(7)[::-1] + 10
                pts[:, -1] = rng.randint(0, 2, 7)
            out["in{}".format(width)] = pts
            out["out{}".format(width)] = sim._EPIFMSimulator__format_inputs(pts)[0][1]
    return out


def main():
    ref = ref_shim.import_reference()
    os.makedirs(OUT, exist_ok=True)
    config = ref.DefaultConfiguration()
    config.default.detector.exposure_time = 33.0e-3

    known = dict(psf=psf_known_answers(ref), scalars=scalar_known_answers(ref, config))

    numpy.savez_compressed(os.path.join(OUT, "radial_profiles.npz"), **radial_profiles(ref, config))

    # C1: examples/tirf.py inputs, expectation only (the EMCCD draw takes 92 s per frame)
    pl = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pl * 0.5
    rng = numpy.random.RandomState(123)
    inputs = rng.uniform(-L_2, +L_2, size=(100, 2))
    expected, data = expectation_case(ref, config, inputs, 0.033)
    known["tirf_c1"] = dict(photons_sum=float(expected.sum()), photons_max=float(expected.max()),
                            expectation_sum=float((0.92 * (expected + 0.01)).sum()),
                            expectation_max=float((0.92 * (expected + 0.01)).max()))
    numpy.savez_compressed(os.path.join(OUT, "tirf_c1.npz"), inputs=inputs,
                           photons=expected.astype(numpy.float64))

    # spots straddling the image border / far outside, deep and beyond the depth cut-off
    small = ref.DefaultConfiguration()
    small.update("default: {detector: {image_size: [96, 80], exposure_time: 0.033}}")
    Lw, Lh = 96 * pl * 0.5, 80 * pl * 0.5
    rng = numpy.random.RandomState(11)
    pts = numpy.zeros((14, 3))
    pts[:, 0] = rng.uniform(-Lw * 1.2, Lw * 1.2, 14)
    pts[:, 1] = rng.uniform(-Lh * 1.2, Lh * 1.2, 14)
    pts[:, 2] = [0, 1e-9, 0.29e-6, 0.5e-6, 0.8e-6, 0.9995e-6, 1.0005e-6, 1.2e-6, 3e-6, 0.1e-6, -0.2e-6, 0, 0, 0]
    pts[11, :2] = [Lw + 1.0e-6, 0.0]      # just outside: still touches the last rows
    pts[12, :2] = [-Lw - 0.9e-6, -Lh - 0.9e-6]
    pts[13, :2] = [5 * Lw, 0.0]           # far outside: no pixel
    expected, data = expectation_case(ref, small, pts, 0.033)
    numpy.savez_compressed(os.path.join(OUT, "border_depth_case.npz"), inputs=pts, formatted=data,
                           photons=expected)

    # Gaussian PSF
    gauss = ref.DefaultConfiguration()
    gauss.update("""
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    detector: {image_size: [64, 64], exposure_time: 0.033}
""")
    rng = numpy.random.RandomState(12)
    pts = rng.uniform(-64 * pl * 0.5, 64 * pl * 0.5, size=(9, 2))
    expected, data = expectation_case(ref, gauss, pts, 0.033)
    numpy.savez_compressed(os.path.join(OUT, "gaussian_case.npz"), inputs=pts, photons=expected)

    # movies through the public API: motion blur (3 snapshots per frame), bleaching, FPN
    numpy.savez_compressed(os.path.join(OUT, "movie_ccd.npz"), **movie_case(ref, """
default:
    detector: {type: CCD, image_size: [24, 20], exposure_time: 0.033, readout_noise: 3.0}
    analog_to_digital_converter: {type: column, count: 2.0}
    effects: {photo_bleaching: {half_life: {value: 0.05, units: s}}}
""", n=12, frames=4, ndim=2, seed=7, size=(24, 20)))
    numpy.savez_compressed(os.path.join(OUT, "movie_cmos3d.npz"), **movie_case(ref, """
default:
    detector: {type: CMOS, image_size: [24, 20], exposure_time: 0.033, QE: 0.73, pixel_length: {value: 6.5e-6, units: m}}
    magnification: 100
    light_source: {angle: {value: 0.0, units: radian}}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: pixel, count: 2.0}
""", n=10, frames=2, ndim=3, seed=9, size=(24, 20)))

    numpy.savez_compressed(os.path.join(OUT, "emccd_pmf.npz"), **emccd_pmfs(ref))
    numpy.savez_compressed(os.path.join(OUT, "format_data.npz"), **format_cases(ref))

    # ADC known answers, _epifm.py:1472-1484
    from scopyon import _epifm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ec = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    adc_fn = _epifm._EPIFMSimulator._EPIFMSimulator__get_analog_to_digital_converter_counts
    pe = numpy.array([0.0, 1.0, 1e6, -5000.0, 123.4, 799999.9, 800000.1])
    known["adc"] = dict(pe=pe.tolist(), counts=adc_fn(
        pe, fullwell=ec.ADConverter_fullwell, gain=ec.ADConverter_gain.ravel()[:len(pe)],
        offset=ec.ADConverter_offset.ravel()[:len(pe)], bit=ec.ADConverter_bit).tolist())

    with open(os.path.join(OUT, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1, sort_keys=True)
    for name in sorted(os.listdir(OUT)):
        print("{:28s} {:9d} bytes".format(name, os.path.getsize(os.path.join(OUT, name))))


if __name__ == "__main__":
    main()
