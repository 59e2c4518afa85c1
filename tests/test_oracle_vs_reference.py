"""The oracle against the LIVE, unmodified reference (only where /root/reference exists,
i.e. in the build container).  This is what pins the oracle; tests/golden/ carries the
same evidence to machines without the reference."""
import warnings

import numpy
import pytest

import epifm_oracle as orc
import ref_shim
from conftest import cmos_table, format_inputs, make_configs

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    warnings.simplefilter("ignore")
    return ref_shim.import_reference()


def test_psf_pieces_bit_for_bit(ref):
    from scopyon import _epifm as R
    rc = R.EPIFMConfigs(ref.DefaultConfiguration().default, rng=numpy.random.RandomState(5))
    _, _, params = make_configs()
    r = orc.radial_grid(1000e-9)
    for z in (0.0, 100e-9, 1000e-9):
        assert numpy.array_equal(rc.fluorophore_psf.get_distribution(r, z),
                                 orc.born_wolf_radial(r, z, params["psf_wavelength"]))
    psf = orc.PsfTables(params)
    for depth in (0.0, 0.29e-6, 2e-6):
        assert numpy.array_equal(psf.get(depth)[1], rc.fluorophore_psf.get(depth))
    rng = numpy.random.RandomState(1)
    table = psf.get(0.0)[1]
    pl = 16e-6 / 241
    for _ in range(4):
        a, b = numpy.zeros((48, 40)), numpy.zeros((48, 40))
        p = numpy.array([0.0, rng.uniform(-2e-6, 2e-6), rng.uniform(-1.6e-6, 1.6e-6)])
        R.PointSpreadingFunction.overlay_signal_(a, table, p, pl, 1e-9, 3.3)
        orc.overlay_signal_exact(b, table, p, pl, 3.3)
        assert numpy.array_equal(a, b)


def test_config_flatten_equals_reference_configs(ref):
    from scopyon import _epifm as R
    yaml = """
default:
    magnification: 100
    detector: {type: CMOS, image_size: [32, 24], QE: 0.73}
    analog_to_digital_converter: {type: pixel, count: 2.0, offset: 100, fullwell: 30000}
"""
    cfg = ref.DefaultConfiguration()
    cfg.update(yaml)
    rc = R.EPIFMConfigs(cfg.default, rng=numpy.random.RandomState(5))
    _, pc, _ = make_configs(yaml)
    sim = R._EPIFMSimulator(rc)
    assert sim.snells_law() == pc.snells_law()
    for name in ("psf_wavelength", "psf_normalization", "image_magnification", "detector_type",
                 "detector_pixel_length", "detector_qeff", "ADConverter_bit", "ADConverter_fullwell",
                 "ADConverter_fpn_type", "ADConverter_fpn_count"):
        assert getattr(rc, name) == getattr(pc, name), name
    assert rc.fluoem_norm.sum() == pc.fluoem_norm_sum
    assert tuple(rc.detector_image_size) == tuple(pc.detector_image_size)


@pytest.mark.parametrize("yaml,ndim,n,frames,size", [
    ("""
default:
    detector: {type: CCD, image_size: [24, 20], exposure_time: 0.033, readout_noise: 3.0}
    analog_to_digital_converter: {type: column, count: 2.0}
""", 2, 12, 3, (24, 20)),
    ("""
default:
    detector: {type: EMCCD, image_size: [10, 8], exposure_time: 0.033}
""", 2, 4, 2, (10, 8)),
])
def test_full_frames_bit_for_bit(ref, yaml, ndim, n, frames, size):
    cfg = ref.DefaultConfiguration()
    cfg.update(yaml)
    config, _, params = make_configs(yaml)
    pl = cfg.default.detector.pixel_length / cfg.default.magnification
    L = size[0] * pl * 0.5
    t = numpy.arange(0, (frames + 1) * 0.033, 0.011)
    inputs = ref.sample_inputs(t, N=n, lower=-L, upper=L, ndim=ndim, D=1e-13, rng=numpy.random.RandomState(7))
    out = list(ref.generate_images(inputs, num_frames=frames, config=cfg, rng=numpy.random.RandomState(8),
                                   full_output=True))
    rng = numpy.random.RandomState(8)
    Nw, Nh = params["image_size"]
    normals = rng.normal(params["adc_offset"], params["fpn_count"], Nh) if params["fpn_type"] == "column" else None
    data = format_inputs(config, inputs)
    states = {}
    psf = orc.PsfTables(params)
    for f in range(frames):
        camera, true_data = orc.output_frame(data, params, rng, frame_index=f, fluorescence_states=states,
                                             cmos_table=cmos_table(), adc_normals=normals, psf=psf, exact=True)
        img, info = out[f]
        assert numpy.array_equal(camera[:, :, 0], info["expectation"])
        assert numpy.array_equal(camera[:, :, 1], img.as_array())
        assert states == info["fluorescence_states"]
        assert all(numpy.array_equal(true_data[k], info["true_data"][k]) for k in true_data)


def test_move_points_same_stream(ref):
    from scopyon import sampling as S
    pts, _ = S.sample_points(numpy.random.RandomState(3), N=50, lower=0, upper=1e-5, ndim=3)
    a = S.move_points(numpy.random.RandomState(4), pts, D=[1e-13, 2e-13, 0.0], dt=0.033, ndim=3)
    b = orc.move_points(numpy.random.RandomState(4), pts, D=[1e-13, 2e-13, 0.0], dt=0.033, ndim=3)
    assert numpy.allclose(a, b, rtol=1e-15, atol=0)


def test_sampling2_bit_for_bit(ref, monkeypatch):
    """oracle/sampling2_oracle.py against the live sampling2.py on the same RandomState: placement,
    per-state Brownian step with periodic wrap, the whole `sample` loop -- and the transition step,
    which the reference can only run once `side='leff'` (sampling2.py:66) is read as 'left'."""
    import sampling2_oracle as s2
    from scopyon import sampling2 as R
    lower, upper = numpy.array([0.0, -1e-6, 2e-7]), numpy.array([1e-6, 1e-6, 2e-7])     # one degenerate axis
    a = getattr(R, "__generate_points")(numpy.random.RandomState(1), N=[30, 0, 12], lower=lower, upper=upper, ndim=3)
    b = s2.generate_points(numpy.random.RandomState(1), [30, 0, 12], lower, upper, 3)
    assert numpy.array_equal(a, b) and a.shape == (42, 5)
    D = numpy.array([1e-12, 0.0, 3e-13])
    for periodic in (False, True):
        ma = getattr(R, "__move_points")(numpy.random.RandomState(2), a, D=D, lower=lower, upper=upper, dt=0.05, ndim=3,
                                         periodic=periodic)
        mb = s2.move_points(numpy.random.RandomState(2), b, D, lower, upper, 0.05, 3, periodic)
        assert numpy.array_equal(ma, mb)
    assert (mb[:, 0] >= 0).all() and (mb[:, 0] < 1e-6).all() and (mb[:, 2] == 2e-7).all()
    transmat = numpy.array([[0.0, 2.0, 0.5], [1.0, 0.0, 0.0], [0.3, 4.0, 0.0]])
    with pytest.raises(ValueError):          # the reference as written
        getattr(R, "__transition_states")(numpy.random.RandomState(3), a, transmat=transmat, dt=0.1, ndim=3)
    real = numpy.searchsorted
    monkeypatch.setattr(numpy, "searchsorted",
                        lambda arr, v, side='left', sorter=None: real(arr, v, side='left' if side == 'leff' else side, sorter=sorter))
    ta = getattr(R, "__transition_states")(numpy.random.RandomState(3), a, transmat=transmat, dt=0.1, ndim=3)
    tb = s2.transition_states(numpy.random.RandomState(3), b, transmat, 0.1, 3)
    assert numpy.array_equal(ta, tb) and len(numpy.unique(tb[:, 3])) == 3
    t = [0.0, 0.1, 0.1, 0.25]
    sa = R.sample(t, [20, 10, 5], lower=lower, upper=upper, D=D, transmat=transmat, ndim=3, periodic=True,
                  rng=numpy.random.RandomState(4))
    sb = s2.sample(t, [20, 10, 5], lower, upper, D, transmat=transmat, ndim=3, periodic=True,
                   rng=numpy.random.RandomState(4))
    assert len(sa) == len(sb) == 4 and all(numpy.array_equal(x, y) for x, y in zip(sa, sb))
