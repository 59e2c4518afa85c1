"""Spot detector: the oracle against the reference's golden output and against
the live reference, and the host-side pieces of scopyon_b200.analysis (no GPU needed)."""
import importlib
import warnings

import numpy
import pytest
import scipy.ndimage

import spot_oracle
from conftest import golden
host = importlib.import_module("scopyon_b200.analysis.spot_detection")    # the package re-exports the function under this name


def _image():
    case = golden("spots_case.npz")
    return case, case["image"].astype(numpy.float64)


def test_oracle_fit_reproduces_the_reference_golden_spots():
    case, image = _image()
    for name, roi in (("roi6", 6), ("roi4p5", 4.5)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = spot_oracle.spot_detection(image, roi_size=roi, blobs=case["blobs"])
            edge = spot_oracle.spot_detection(image, roi_size=roi, blobs=case["edge_blobs"])
        assert got.shape == case["spots_" + name].shape
        numpy.testing.assert_allclose(got, case["spots_" + name], rtol=1e-9, atol=1e-9)
        assert edge.shape == case["edge_spots_" + name].shape
        numpy.testing.assert_allclose(edge, case["edge_spots_" + name], rtol=1e-9, atol=1e-9)


def test_oracle_blobs_are_the_golden_blobs():
    case, image = _image()
    blobs = spot_oracle.blob_detection(image, min_sigma=1, max_sigma=4, threshold=50.0, overlap=0.5)
    numpy.testing.assert_allclose(blobs, case["blobs"], rtol=0, atol=1e-12)


@pytest.mark.reference
def test_oracle_fit_equals_live_reference():
    import ref_shim
    warnings.simplefilter("ignore")
    ref = ref_shim.import_reference()
    case, image = _image()
    rng = numpy.random.RandomState(4)
    blobs = numpy.column_stack([rng.uniform(0, 191, 25), rng.uniform(0, 191, 25), numpy.full(25, 1.4)])
    blobs = numpy.concatenate([case["blobs"], blobs])
    for roi in (6, 3, 7.5):
        assert numpy.array_equal(spot_oracle.spot_detection(image, roi_size=roi, blobs=blobs),
                                 ref.analysis.spot_detection(image, roi_size=roi, blobs=blobs))


def test_blob_log_known_answer_single_gaussian():
    """One Gaussian bump of width 3: the scale-normalised LoG peaks at its centre with sigma = 3,
    value = amplitude / 2 for an exact Gaussian."""
    X, Y = numpy.indices((64, 80))
    image = 100.0 * numpy.exp(-((X - 20) ** 2 + (Y - 31) ** 2) / (2 * 3.0 ** 2))
    blobs = spot_oracle.blob_log(image, min_sigma=1, max_sigma=5, num_sigma=5, threshold=10.0)
    assert blobs.tolist() == [[20.0, 31.0, 3.0]]
    cube = spot_oracle.log_cube(image, spot_oracle.sigma_list(1, 5, 5))
    assert abs(cube[20, 31, 2] - 50.0) < 1.0
    assert spot_oracle.blob_log(numpy.zeros((16, 16)), max_sigma=3, threshold=-1.0).shape == (0, 3)    # flat cube: no peaks


def test_host_half_kernels_are_scipys_filter_taps():
    for sigma in (1.0, 1.75, 4.0, 9.3):
        radius, g0, g2 = host.gaussian_half_kernels(sigma)
        assert radius == int(4.0 * sigma + 0.5)
        delta = numpy.zeros(2 * radius + 1)
        delta[radius] = 1.0
        for order, half in ((0, g0), (2, g2)):
            taps = scipy.ndimage.gaussian_filter1d(delta, sigma, order=order, mode='constant')
            assert numpy.array_equal(taps[radius:], half)
            assert numpy.array_equal(taps[: radius + 1][::-1], half)


def test_host_prune_matches_oracle_prune():
    rng = numpy.random.RandomState(6)
    for trial in range(20):
        n = rng.randint(2, 60)
        blobs = numpy.column_stack([rng.uniform(0, 40, n).round(), rng.uniform(0, 40, n).round(),
                                    rng.choice([1.0, 1.75, 2.5, 3.25, 4.0], n)])
        for overlap in (0.1, 0.5, 0.9):
            want = spot_oracle.prune_blobs(blobs.copy(), overlap)
            got = host.prune_blobs(blobs.copy(), overlap)
            assert numpy.array_equal(got, want)


def test_host_image_conversion_follows_img_as_float():
    assert host._as_float_image(numpy.array([[0, 255]], dtype=numpy.uint8)).tolist() == [[0.0, 1.0]]
    assert host._as_float_image(numpy.array([[0, 65535]], dtype=numpy.uint16)).tolist() == [[0.0, 1.0]]
    f = numpy.array([[1.5, -2.0]], dtype=numpy.float32)
    assert host._as_float_image(f).dtype == numpy.float64 and host._as_float_image(f).tolist() == [[1.5, -2.0]]
    with pytest.raises(TypeError):
        host._as_float_image(numpy.array([[1, 2]], dtype=numpy.int32))
    with pytest.raises(ValueError):
        host._as_float_image(numpy.zeros((2, 2, 2)))
