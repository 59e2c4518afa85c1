"""Timings of the UNMODIFIED reference, taken in the build container
where /root/reference lives: C1 (examples/tirf.py) end to end on one core, the per-particle path at
N = 2000 (overlay ms per spot), move_points per particle-step, detector microseconds per pixel.

    python tools/time_reference.py [--quick]      ->  profiles/reference_live_r2.json

/root/reference does not travel to the GPU box, so bench.py carries this record along
(`reference_live` of the `--impl reference` line), labelled with the machine it was measured on.
The reference is imported through oracle/ref_shim.py (stand-ins for pint / hmmlearn only)."""
import json
import os
import platform
import sys
import time
import warnings

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def main():
    quick = "--quick" in sys.argv
    warnings.simplefilter("ignore")
    scopyon = ref_shim.import_reference()
    from scopyon import _epifm as R
    from scopyon import sampling as S
    import scipy
    out = {"measured_on": {"where": "build container (no GPU), NOT the B200 box", "cpu": cpu_model(),
                           "cores_visible": os.cpu_count(), "python": platform.python_version(),
                           "numpy": numpy.__version__, "scipy": scipy.__version__},
           "reference": "ecell/scopyon, unmodified, imported from /root/reference/src through oracle/ref_shim.py"}

    # ---- C1: examples/tirf.py body, 512^2 EMCCD, 100 spots, 1 frame, one core
    if not quick:
        config = scopyon.DefaultConfiguration()
        config.default.detector.exposure_time = 33.0e-3
        pl = config.default.detector.pixel_length / config.default.magnification
        L_2 = config.default.detector.image_size[0] * pl * 0.5
        rng = numpy.random.RandomState(123)
        inputs = rng.uniform(-L_2, +L_2, size=(100, 2))
        t0 = time.perf_counter()
        scopyon.form_image(inputs, config=config, rng=rng)
        dt = time.perf_counter() - t0
        out["c1_tirf_form_image"] = {"seconds": dt, "frames_per_s": 1.0 / dt, "cores": 1,
                                     "what": "examples/tirf.py body: 512x512 EMCCD, 100 molecules, 1 frame"}

    # ---- per-particle path (emission + overlay), N = 2000, CCD 512^2, 2-D, one core
    n = 400 if quick else 2000
    config = scopyon.DefaultConfiguration()
    config.default.detector.type = "CCD"
    config.default.detector.exposure_time = 33.0e-3
    pl = config.default.detector.pixel_length / config.default.magnification
    L_2 = config.default.detector.image_size[0] * pl * 0.5
    rng = numpy.random.RandomState(5)
    inputs = rng.uniform(-L_2, +L_2, size=(n, 2))
    sim = scopyon.create_simulator(config, rng=rng)
    base = sim.base()
    data = ((0.0, sim._EPIFMSimulator__format_data(inputs)),)
    cfg = base.configs
    p_b = numpy.array([0.0, 0.0, 0.0])
    p_0 = numpy.asarray(cfg.detector_focal_point) if hasattr(cfg, "detector_focal_point") else p_b
    shape = tuple(config.default.detector.image_size)
    t0 = time.perf_counter()
    expected, _, _ = base.get_molecule_plane(data[0][1], shape, p_b, p_0, 33.0e-3, {}, None, rng=rng)
    dt = time.perf_counter() - t0
    footprint = float((numpy.ceil(1998e-9 / pl) + 1) ** 2)
    out["overlay_per_particle_path"] = {
        "seconds": dt, "spots": n, "ms_per_spot": dt / n * 1e3, "cores": 1,
        "spot_pixel_evals_per_s": n * footprint / dt,
        "what": "_EPIFMSimulator.get_molecule_plane (emission + overlay_signal_), CCD 512x512, 2-D, {:.2f} nm pixels".format(pl * 1e9)}

    # ---- move_points, microseconds per particle-step
    n_mp = 2000 if quick else 10000
    pts, _ = S.sample_points(numpy.random.RandomState(3), N=n_mp, lower=0, upper=1e-5, ndim=3)
    t0 = time.perf_counter()
    S.move_points(numpy.random.RandomState(4), pts, D=1e-13, dt=0.033, ndim=3)
    dt = time.perf_counter() - t0
    out["move_points"] = {"seconds": dt, "particles": n_mp, "us_per_particle_step": dt / n_mp * 1e6, "cores": 1}

    # ---- detector, microseconds per pixel
    rng = numpy.random.RandomState(9)
    side = 64 if quick else 256
    expected = rng.gamma(2.0, 0.5, (side, side))
    t0 = time.perf_counter()
    R.CMOS.get_noise((side, side), rng)
    t_noise = time.perf_counter() - t0
    t0 = time.perf_counter()
    R.CMOS.get_signal(expected, rng=rng)
    t_signal = time.perf_counter() - t0
    out["cmos"] = {"pixels": side * side, "get_noise_us_per_pixel": t_noise / side ** 2 * 1e6,
                   "get_signal_us_per_pixel": t_signal / side ** 2 * 1e6, "cores": 1}
    t0 = time.perf_counter()
    R.CCD.get_signal(expected, rng=rng)
    out["ccd"] = {"pixels": side * side, "get_signal_us_per_pixel": (time.perf_counter() - t0) / side ** 2 * 1e6, "cores": 1}
    side_em = 16 if quick else 64
    expected = rng.gamma(2.0, 0.5, (side_em, side_em))
    t0 = time.perf_counter()
    R.EMCCD.get_signal(expected, 300.0, rng=rng)
    out["emccd"] = {"pixels": side_em ** 2, "get_signal_us_per_pixel": (time.perf_counter() - t0) / side_em ** 2 * 1e6,
                    "cores": 1}

    # ---- C4 extrapolated from the stage costs (linear: overlay ~ spots, detector ~ pixels), one core
    frame_s = 1e5 * (out["overlay_per_particle_path"]["ms_per_spot"] * 1e-3 + out["move_points"]["us_per_particle_step"] * 1e-6) \
        + 2048 * 2048 * (out["cmos"]["get_noise_us_per_pixel"] + out["cmos"]["get_signal_us_per_pixel"]) * 1e-6
    out["c4_extrapolated"] = {"seconds_per_frame": frame_s, "frames_per_s": 1.0 / frame_s, "cores": 1,
                              "how": "1e5 x (overlay + move_points per spot) + 2048^2 x CMOS (noise + signal) per pixel; "
                                     "2-D overlay cost (a 3-D movie adds one 0.12 s PSF-table build per new depth key)"}
    path = os.path.join(ROOT, "profiles", "reference_live_r2.json")
    if not quick:
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
