"""The CPU oracle against the golden vectors made from the live reference
(oracle/make_golden.py) -- runs anywhere, pins the oracle on the GPU box too."""
import numpy
import pytest

import epifm_oracle as orc
from conftest import cmos_table, format_inputs, golden, make_configs


def test_reference_test_suite_integrals(known_answers):
    # the six numbers the reference's own test prints (test/test_epifm.py:30-42, 58-70)
    radial = orc.radial_grid(1000.0e-9)
    for name, prof in (("tritc", orc.born_wolf_radial(radial, 0.0, 5.78e-07)),
                       ("gaussian", orc.gaussian_radial(radial, 1.0e-7))):
        want = known_answers["psf"][name]
        assert numpy.sum(2 * numpy.pi * radial * prof) * 1.0e-9 == want["radial_integral"]
        cart = orc.radial_to_cartesian(radial, prof)
        assert cart.sum() * 1.0e-18 == want["cartesian_integral"]
        assert cart[999, 999] == want["table_centre"] and cart[0, 0] == want["table_corner"]
        camera = numpy.zeros((512, 512))
        orc.overlay_signal_exact(camera, cart, numpy.zeros(3), 4.444444444444444e-08, 1.0)
        assert camera.sum() == want["overlay_sum"]
        assert (camera > 0).sum() == want["footprint_pixels"]


def test_scalar_known_answers(known_answers):
    _, _, params = make_configs("default: {detector: {exposure_time: 0.033}}")
    want = known_answers["scalars"]
    amplitude, depth = orc.snells_law(params)
    assert amplitude == want["snells_amplitude"] and depth == want["snells_depth"]
    assert orc.emit_photons(amplitude, 0.033, 83400, 0.61, 20e-9) == want["n_emit_33ms"]
    assert params["psf_wavelength"] == want["psf_wavelength"]
    assert params["fluoem_norm_sum"] == want["fluoem_norm_sum"]
    assert orc.HC == want["hc"] and orc.N_A == want["N_A"]
    offset, gain = orc.adc_params(params)
    assert gain[0, 0] == want["adc_gain_none"]
    adc = known_answers["adc"]
    n = len(adc["pe"])
    got = orc.adc_counts(adc["pe"], params["adc_fullwell"], gain.ravel()[:n], offset.ravel()[:n], params["adc_bit"])
    assert got.tolist() == adc["counts"]


def test_radial_profiles_and_depth_keys():
    g = golden("radial_profiles.npz")
    radial = orc.radial_grid(1000.0e-9)
    for z, want in zip(g["depths"], g["born_wolf"]):
        assert numpy.array_equal(orc.born_wolf_radial(radial, float(z), float(g["psf_wavelength"])), want)
    # _epifm.py:76-84; 0.29e-6 -> key 289 is the fp64 edge case noted in SURVEY.md
    assert orc.depth_key(0.29e-6, 1000e-9) == (289, 289 * 1e-9)
    assert orc.depth_key(-5.5e-9, 1000e-9)[0] == 5
    assert orc.depth_key(1.0009e-6, 1000e-9)[0] == 1000
    assert orc.depth_key(1.0011e-6, 1000e-9) == (-1, 1000e-9)


def test_table_integrals_by_depth():
    g = golden("radial_profiles.npz")
    _, _, params = make_configs()
    psf = orc.PsfTables(params)
    for z, want in zip((0.0, 100e-9, 500e-9, 1000e-9), g["table_integrals"]):
        assert psf.get(z)[1].sum() * 1e-18 == want


@pytest.mark.parametrize("exact", [True, False])
def test_tirf_c1_expectation(known_answers, exact):
    """examples/tirf.py (BASELINE config 1): pre-noise image of the 100 seeded spots."""
    g = golden("tirf_c1.npz")
    config, configs, params = make_configs("default: {detector: {exposure_time: 0.033}}")
    data = format_inputs(config, g["inputs"])
    photons, true_data = orc.expected_frame(data, params, exact=exact)
    if exact:
        assert numpy.array_equal(photons, g["photons"])
        assert photons.sum() == known_answers["tirf_c1"]["photons_sum"]
        expectation = orc.detector_expectation(photons, params)
        assert expectation.sum() == known_answers["tirf_c1"]["expectation_sum"]
        assert expectation.max() == known_answers["tirf_c1"]["expectation_max"]
    else:
        assert abs(photons - g["photons"]).max() / g["photons"].max() < 1e-12
    assert len(true_data) == 100 and true_data[0][0] == 0.033


def test_border_depth_and_gaussian_cases():
    g = golden("border_depth_case.npz")
    config, _, params = make_configs("default: {detector: {image_size: [96, 80], exposure_time: 0.033}}")
    data = format_inputs(config, g["inputs"])
    assert numpy.array_equal(data[0][1], g["formatted"])
    photons, _ = orc.expected_frame(data, params, exact=True)
    assert numpy.array_equal(photons, g["photons"])

    g = golden("gaussian_case.npz")
    config, _, params = make_configs("""
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    detector: {image_size: [64, 64], exposure_time: 0.033}
""")
    photons, _ = orc.expected_frame(format_inputs(config, g["inputs"]), params, exact=True)
    assert numpy.array_equal(photons, g["photons"])


@pytest.mark.parametrize("name", ["movie_ccd.npz", "movie_cmos3d.npz"])
def test_movies_bit_for_bit(name):
    """Full frames (motion blur, bleaching, FPN, detector, ADC) with the reference's
    RandomState draw order reproduce the reference's output exactly."""
    g = golden(name)
    config, _, params = make_configs(str(g["yaml"]))
    inputs = [(float(t), p) for t, p in zip(g["times"], g["points"])]
    data = format_inputs(config, inputs)
    rng = numpy.random.RandomState(int(g["seed"]) + 1)
    Nw, Nh = params["image_size"]
    normals = None
    if params["fpn_type"] == "pixel":
        normals = rng.normal(params["adc_offset"], params["fpn_count"], Nw * Nh)
    elif params["fpn_type"] == "column":
        normals = rng.normal(params["adc_offset"], params["fpn_count"], Nh)
    states = {}
    psf = orc.PsfTables(params)
    for f in range(g["adc"].shape[0]):
        camera, true_data = orc.output_frame(
            data, params, rng, frame_index=f, fluorescence_states=states, cmos_table=cmos_table(),
            adc_normals=normals, psf=psf, exact=True)
        assert numpy.array_equal(camera[:, :, 0], g["expectation"][f])
        assert numpy.array_equal(camera[:, :, 1], g["adc"][f])
        for i, want in zip(g["true_ids"], g["true_data"][f]):
            assert numpy.array_equal(true_data[int(i)], want)
        assert [states[int(i)] for i in g["true_ids"]] == g["budgets"][f].tolist()


def test_emccd_pmf_matches_reference():
    g = golden("emccd_pmf.npz")
    for i in range(4):
        E = float(g["E{}".format(i)])
        S, p = orc.emccd_pmf(E, 300)
        cdf = numpy.cumsum(p)
        step = max(1, len(S) // 2000)
        assert numpy.array_equal(S[::step], g["S{}".format(i)])
        assert numpy.array_equal(cdf[::step], g["cdf{}".format(i)])
        assert (S * p).sum() == float(g["mean{}".format(i)])


def test_sampling2_oracle_against_reference_golden():
    """oracle/sampling2_oracle.py reproduces the reference's multi-state trajectories
    (tests/golden/sampling2_case.npz, made by oracle/make_golden_sampling2.py from the live reference)
    bit for bit on the same RandomState."""
    import sampling2_oracle as s2
    g = golden("sampling2_case.npz")
    t = list(g["t"])
    out = s2.sample(t, [40, 25, 10], g["lower"], g["upper"], g["D"], transmat=g["transmat"], ndim=3, periodic=True,
                    rng=numpy.random.RandomState(2024))
    assert numpy.array_equal(numpy.stack(out), g["periodic_switching"])
    free = s2.sample(t, [30, 30], numpy.zeros(2), numpy.ones(2) * 1e-6, numpy.array([2e-13, 5e-12]), ndim=2,
                     periodic=False, rng=numpy.random.RandomState(7))
    assert numpy.array_equal(numpy.stack(free), g["free"])
    assert len(numpy.unique(g["periodic_switching"][-1][:, 3])) == 3          # states did switch
