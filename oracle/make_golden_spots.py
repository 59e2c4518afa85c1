"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the spot detector, made by the LIVE reference
(``/root/reference/src/scopyon/analysis/spot_detection.py``) in the build container:

    python oracle/make_golden_spots.py      ->  tests/golden/spots_case.npz

The camera image is the C1 expectation (``tests/golden/tirf_c1.npz``, examples/tirf.py) cropped to
192 x 192, pushed through an EMCCD-like draw (Poisson -> gamma gain 300 -> read noise 100 e- -> ADC)
with a fixed numpy seed and stored as float32.  The blobs come from the restated ``blob_log``
(oracle/spot_oracle.py; scikit-image is not installed, so that half is unpinned); the spots are
what the reference's ``spot_detection(data, roi_size, blobs=blobs)`` returns for them, for two ROI
sizes, plus a set of blobs at and beyond the image border.
"""
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim          # noqa: E402
import spot_oracle       # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def camera_image():
    photons = numpy.load(os.path.join(OUT, "tirf_c1.npz"))["photons"][100:292, 60:252]
    rng = numpy.random.RandomState(3)
    expected = 0.92 * (photons + 0.01)
    n = rng.poisson(expected)
    electrons = numpy.where(n > 0, rng.gamma(numpy.maximum(n, 1), 300.0), 0.0) + rng.normal(0, 100, expected.shape)
    gain = 800000 / (65536 - 2000)
    adc = numpy.clip(numpy.minimum(electrons, 800000) / gain + 2000, 0, 65535)
    return adc.astype(numpy.float32).astype(numpy.float64)


def main():
    warnings.simplefilter("ignore")
    ref = ref_shim.import_reference()
    image = camera_image()
    blobs = spot_oracle.blob_detection(image, min_sigma=1, max_sigma=4, threshold=50.0, overlap=0.5)
    rng = numpy.random.RandomState(8)
    edge = numpy.array([[0.0, 0.0, 1.4], [191.0, 191.0, 1.4], [0.0, 100.3, 1.4], [95.5, 191.0, 2.0],
                        [3.2, 4.9, 1.4], [188.7, 2.1, 1.4]])
    edge = numpy.concatenate([edge, numpy.column_stack([rng.uniform(0, 191, 6), rng.uniform(0, 191, 6),
                                                        numpy.full(6, 1.4)])])
    out = dict(image=image.astype(numpy.float32), blobs=blobs, edge_blobs=edge)
    for name, roi in (("roi6", 6), ("roi4p5", 4.5)):
        out["spots_" + name] = ref.analysis.spot_detection(image, roi_size=roi, blobs=blobs)
        out["edge_spots_" + name] = ref.analysis.spot_detection(image, roi_size=roi, blobs=edge)
    numpy.savez_compressed(os.path.join(OUT, "spots_case.npz"), **out)
    for k, v in out.items():
        print(k, v.shape)


if __name__ == "__main__":
    main()
