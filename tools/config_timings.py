"""Wall-clock of BASELINE.json configs 1-3 through the public API (examples/tirf.py, twocolor.py,
bleaching.py of the reference, run with `import scopyon_b200 as scopyon`).  Prints one JSON line
per config: first call (PSF tables built, buffers allocated) and steady state."""
import json
import os
import sys
import time
import warnings

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scopyon_b200 as scopyon  # noqa: E402

warnings.simplefilter("ignore")


def timed(fn, repeat=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return first, float(numpy.median(times))


def c1():
    """examples/tirf.py: 100 molecules, 512^2 EMCCD, one 33 ms frame."""
    config = scopyon.DefaultConfiguration()
    config.default.detector.exposure_time = 33.0e-3
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    rng = numpy.random.RandomState(123)
    inputs = rng.uniform(-L_2, +L_2, size=(100, 2))
    return lambda: scopyon.form_image(inputs, config=config, rng=rng)


def c2():
    """examples/twocolor.py: two channels (200 and 100 molecules) and Image.RGB."""
    config = scopyon.DefaultConfiguration()
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    rng = numpy.random.RandomState(123)
    inputs = rng.uniform(-L_2, +L_2, size=(250, 2))

    def run():
        img1 = scopyon.form_image(inputs[: 200], config=config, rng=rng)
        img2 = scopyon.form_image(inputs[150:], config=config, rng=rng)
        return scopyon.Image.RGB(red=img1, green=img2)
    return run


def c3(frames):
    """examples/bleaching.py: 1000 molecules diffusing in 2-D, x360, photobleaching, `frames` frames."""
    config = scopyon.DefaultConfiguration()
    config.default.magnification = 360
    config.default.detector.exposure_time = 33.0e-3
    config.default.effects.photo_bleaching.half_life = 2.5
    pixel_length = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pixel_length * 0.5
    rng = numpy.random.RandomState(123)
    dt = config.default.detector.exposure_time
    t = numpy.arange(0, (frames + 1) * dt, dt)
    inputs = scopyon.sample_inputs(t, N=1000, lower=-L_2, upper=+L_2, ndim=2, D=0.1e-12, rng=rng)
    return lambda: list(scopyon.generate_images(inputs, num_frames=frames, config=config, rng=rng))


if __name__ == "__main__":
    for name, fn, ref in (("C1 tirf.py", c1(), "reference: 91.9 s (SURVEY.md section 6, one core)"),
                          ("C2 twocolor.py", c2(), "reference: two EMCCD frames, ~2 x 92 s"),
                          ("C3 bleaching.py, 30 frames", c3(30), "reference: 30 EMCCD frames, ~30 x 100 s"),
                          ("C3 bleaching.py, 1000 frames", c3(1000), "")):
        first, steady = timed(fn, repeat=3)
        print(json.dumps({"config": name, "first_call_s": first, "steady_s": steady, "note": ref}))
