"""The reference's notebook "testsuite" (docs/examples/testsuite1.ipynb, cells 3-14) on the GPU: images of
1000 molecules at x360 on the default EMCCD, spot detection, and the distance from every detected spot
to the closest true molecule.  The simulate -> detect loop closes through the public API alone
(form_image, true_data, analysis.spot_detection), so image formation, the EMCCD sampler and the
detector-side fit must all be right for these statistics to come out.

Golden numbers: tests/golden/testsuite_case.json, made by oracle/make_golden_testsuite.py from the LIVE
reference (six images, seeds 123..128, two CPU-minutes each).  The notebook's own printed numbers
(+0.035 / +0.002 px, std 1.13 / 1.14 px, from an older release and scikit-image) are kept in the file
for the record; today's reference gives std 1.33 / 1.29 px with the same procedure."""
import warnings

import numpy
import pytest

import scopyon_b200
from conftest import golden

pytestmark = pytest.mark.gpu


def test_localisation_statistics_match_the_live_reference():
    want = golden("testsuite_case.json")
    config = scopyon_b200.DefaultConfiguration()
    config.update("""
default:
    magnification: 360
    detector:
        exposure_time: 0.033
""")
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    N = 1000
    closest, counts = [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for ref_image in want["images"]:
            rng = numpy.random.RandomState(ref_image["seed"])         # the same molecules as the reference's image
            inputs = rng.uniform(-L_2, +L_2, size=(N, 2))
            img, infodict = scopyon_b200.form_image(inputs, config=config, rng=rng, full_output=True)
            image = img.as_array()
            # same expectation, independent noise: 512^2 pixels of sigma ~ 30 counts
            assert abs(image.mean() - ref_image["image_mean"]) < 6 * ref_image["image_std"] / 512
            assert abs(image.std() - ref_image["image_std"]) < 0.6
            spots = scopyon_b200.analysis.spot_detection(image, min_sigma=1, max_sigma=4, threshold=40.0, overlap=0.5)
            data = numpy.array([(d[2], d[3]) for d in infodict['true_data'].values()])
            assert len(data) == N
            counts.append(len(spots))
            for spot in spots:
                distance = data - spot[0: 2]
                closest.append(distance[(distance ** 2).sum(axis=1).argmin()])
    closest = numpy.array(closest)
    n = len(closest)
    radial = numpy.sqrt((closest ** 2).sum(axis=1))
    ref_counts = [im["spots"] for im in want["images"]]
    print("spots per image", counts, "reference", ref_counts)
    print("mean", closest.mean(axis=0), want["mean"], "std", closest.std(axis=0), want["std"])
    print("within 1 / 2 px, beyond 4 px", (radial < 1).mean(), (radial < 2).mean(), (radial > 4).mean(),
          want["within_1px"], want["within_2px"], want["beyond_4px"])
    assert abs(numpy.mean(counts) / numpy.mean(ref_counts) - 1) < 0.04
    sigma = numpy.mean(want["std"])
    two_samples = numpy.sqrt(1.0 / n + 1.0 / want["n"])
    assert abs(closest.mean(axis=0) - want["mean"]).max() < 5 * sigma * two_samples
    assert abs(closest.std(axis=0) - want["std"]).max() < 0.12        # heavy tails (false detections): ~3 sigma of the two-sample error
    assert abs(numpy.median(abs(closest), axis=0) - want["median_abs"]).max() < 0.03
    for got, ref in (((radial < 1).mean(), want["within_1px"]), ((radial < 2).mean(), want["within_2px"]),
                     ((radial > 4).mean(), want["beyond_4px"])):
        assert abs(got - ref) < 5 * numpy.sqrt(ref * (1 - ref)) * two_samples + 0.005
