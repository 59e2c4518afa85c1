"""TEST INFRASTRUCTURE ONLY.  The reference's notebook "testsuite" (docs/examples/testsuite1.ipynb,
cells 3-14) run by the LIVE reference in the build container:

    python oracle/make_golden_testsuite.py [n_images]     ->  tests/golden/testsuite_case.json

For each image: 1000 molecules at x360 on the default EMCCD through the reference's own
``scopyon.form_image`` (about two minutes per image: the per-pixel EMCCD pmf), then spot detection and
the distance of every detected spot to the closest true molecule (cells 10-12).  scikit-image is not
installed, so the blobs come from the restated ``blob_log`` (oracle/spot_oracle.py) and the fit from the
reference's ``spot_detection(data, blobs=...)``.  The file stores per-image counts and the pooled
distance statistics; tests/test_gpu_testsuite.py compares the GPU pipeline's statistics with them.
"""
import json
import multiprocessing
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "tests", "golden", "testsuite_case.json")
N = 1000


def one_image(seed):
    warnings.simplefilter("ignore")
    import ref_shim
    import spot_oracle
    ref = ref_shim.import_reference()
    config = ref.DefaultConfiguration()
    config.update("""
default:
    magnification: 360
    detector:
        exposure_time: 0.033
""")
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    rng = numpy.random.RandomState(seed)
    inputs = rng.uniform(-L_2, +L_2, size=(N, 2))
    img, infodict = ref.form_image(inputs, config=config, rng=rng, full_output=True)
    image = img.as_array()
    blobs = spot_oracle.blob_detection(image, min_sigma=1, max_sigma=4, threshold=40.0, overlap=0.5)
    spots = ref.analysis.spot_detection(image, blobs=blobs)
    data = numpy.array([(d[2], d[3]) for d in infodict['true_data'].values()])
    closest = []
    for spot in spots:
        distance = data - spot[0: 2]
        closest.append(distance[(distance ** 2).sum(axis=1).argmin()])
    return dict(seed=seed, blobs=int(len(blobs)), spots=int(len(spots)), closest=numpy.array(closest).tolist(),
                image_mean=float(image.mean()), image_std=float(image.std()), image_max=float(image.max()))


def main():
    n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    seeds = [123 + k for k in range(n_images)]
    with multiprocessing.Pool(min(n_images, os.cpu_count() or 1)) as pool:
        results = pool.map(one_image, seeds)
    closest = numpy.concatenate([numpy.array(r["closest"]) for r in results])
    radial = numpy.sqrt((closest ** 2).sum(axis=1))
    summary = dict(
        images=[{k: v for k, v in r.items() if k != "closest"} for r in results],
        n=int(len(closest)), mean=closest.mean(axis=0).tolist(), std=closest.std(axis=0).tolist(),
        median_abs=numpy.median(abs(closest), axis=0).tolist(),
        within_1px=float((radial < 1.0).mean()), within_2px=float((radial < 2.0).mean()),
        beyond_4px=float((radial > 4.0).mean()),
        notebook=dict(mean=[0.03475, 0.00178], std=[1.134670082759238, 1.1447345849828259]),
        note="live reference form_image (seed 123 + k per image) + restated blob_log + reference spot_detection fit")
    with open(OUT, "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "images"}, indent=1))
    print(summary["images"])


if __name__ == "__main__":
    main()
