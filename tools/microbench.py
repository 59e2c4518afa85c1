"""Per-kernel microbenchmarks on one GPU (CUDA events, L2 flushed between launches).

  python tools/microbench.py [detector] [diffuse] [render] [tables]

Prints one JSON object per measurement: achieved GB/s against the algorithmic bytes of
SURVEY.md section 8(d) (8 B per pixel-frame for noise+ADC, 48 B per particle-step fp64
diffusion, 8 B per spot-pixel eval for the render)."""
import ctypes
import json
import os
import sys
import time

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scopyon_b200  # noqa: E402
from scopyon_b200 import _epifm, _native  # noqa: E402
from scopyon_b200.engine import DeviceEngine  # noqa: E402
from scopyon_b200.sampling import DeviceParticles  # noqa: E402

PEAK = 6545.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.fill_(1)


def timed(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(numpy.median(times)), float(min(times))


def engine_for(yaml, precision="f32"):
    import warnings
    config = scopyon_b200.DefaultConfiguration()
    config.update(yaml)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    return configs, DeviceEngine(configs, precision=precision)


def bench_detector():
    for det, extra in (("CMOS", "analog_to_digital_converter: {type: column, count: 2.0, offset: 100, fullwell: 30000}"),
                       ("CMOS", "analog_to_digital_converter: {type: none, offset: 100, fullwell: 30000}"),
                       ("CCD", "analog_to_digital_converter: {type: none}"),
                       ("EMCCD", "analog_to_digital_converter: {type: none}")):
        for size in (2048, 4096):
            yaml = "default:\n    detector: {type: %s, image_size: [%d, %d], QE: 0.73}\n    %s\n" % (det, size, size, extra)
            configs, eng = engine_for(yaml)
            for level, label in ((0.0, "background only"), (0.6, "C4-like 0.6 photons/px"), (20.0, "bright 20 photons/px")):
                photons = torch.full((size, size), level, dtype=torch.float32, device="cuda")
                adc = torch.empty_like(photons)
                ms, best = timed(lambda: eng.detect(photons, 1, 42, adc=adc))
                gbs = size * size * 8 / (ms * 1e-3) / 1e9
                print(json.dumps({"kernel": "detector_kernel<float,%s>" % det, "fpn": configs.ADConverter_fpn_type,
                                  "size": size, "signal": label, "ms": ms, "ms_best": best, "GB/s": gbs,
                                  "frac_of_measured_hbm": gbs / PEAK, "frames/s": 1e3 / ms}))


def bench_detector_block():
    """The detector pass as the C4 step runs it: 16 frames of 2048^2 per launch (scb_detector_adc_frames),
    CMOS + column FPN, C4-like signal.  SCB_DETECTOR_ROUNDS selects the measured generator variants."""
    sys.path.insert(0, ROOT)
    from bench import C4_YAML
    size, nf = 2048, 16
    configs, eng = engine_for(C4_YAML % (size, size, 2.5))
    rng = numpy.random.RandomState(0)
    photons = torch.from_numpy(rng.gamma(0.5, 1.2, (nf, size, size)).astype(numpy.float32)).cuda()   # mean 0.6 photons/px
    adc = torch.empty_like(photons)
    work = torch.empty(nf * eng.lib.scb_detector_workspace_bytes(size, size), dtype=torch.uint8, device="cuda")

    def go():
        eng._call("scb_detector_adc_frames", 42, 0, nf, ctypes.byref(eng.det), size, size, _native.F32,
                  _native.ptr(photons), _native.ptr(eng.offset), _native.ptr(eng.alias), int(eng.alias.shape[0]),
                  _native.ptr(adc), _native.ptr(work), work.numel(), eng._stream())
    ms, best = timed(go, iters=10, warm=3)
    gbs = nf * size * size * 8 / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "detector_fast_kernel<CMOS,column> + slow pass, 16 x 2048^2 per launch",
                      "philox_rounds": int(os.environ.get("SCB_DETECTOR_ROUNDS", "10")), "ms": ms, "ms_best": best,
                      "GB/s": gbs, "frac_of_measured_hbm": gbs / PEAK, "mean_adc": float(adc.mean().item())}))


def bench_diffuse():
    for n in (100000, 10000000, 50000000):
        parts = DeviceParticles(n, 3)
        ms, best = timed(lambda: parts.step(1, 0, [1e-8, 1e-8, 1e-8]))
        gbs = n * 48 / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "diffuse_kernel", "particles": n, "ms": ms, "ms_best": best, "GB/s": gbs,
                          "frac_of_measured_hbm": gbs / PEAK, "particle_steps/s": n / (ms * 1e-3)}))


def bench_render():
    sys.path.insert(0, ROOT)
    from bench import C4_YAML, count_spot_pixel_evals
    for size, n, three_d in ((2048, 100000, False), (2048, 100000, True), (512, 1000, False)):
        configs, eng = engine_for(C4_YAML % (size, size, 2.5))
        pl = configs.pixel_length
        rng = numpy.random.RandomState(1)
        data = numpy.zeros((n, 5))
        data[:, 1:3] = rng.uniform(-size * pl / 2, size * pl / 2, (n, 2))
        if three_d:
            data[:, 0] = rng.uniform(0, 1.5e-6, n)
        data[:, 3] = numpy.arange(n)
        data[:, 4] = 1
        t0 = time.time()
        eng.ensure_all_tables() if three_d else eng.ensure_tables([0])
        torch.cuda.synchronize()
        build_s = time.time() - t0
        soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).cuda()
        w = torch.full((n,), 30.0, dtype=torch.float64, device="cuda")
        out = torch.empty((size, size), dtype=torch.float32, device="cuda")
        work = eng._render_workspace(n)

        def go():
            eng._call("scb_render_expected", ctypes.byref(eng.geom), n, _native.ptr(soa[0]), _native.ptr(soa[1]),
                      _native.ptr(soa[2]), _native.ptr(w), _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type, _native.ptr(eng.inv_scale),
                      _native.ptr(eng.slot_of_key), _native.ptr(out), _native.F32, 0, _native.ptr(work), work.numel(),
                      _native.ptr(eng.errors), eng._stream())
        ms, best = timed(go)
        evals = count_spot_pixel_evals(data, size, pl)
        print(json.dumps({"kernel": "scb_render_expected (prepare+scan+fill+render)", "size": size, "spots": n,
                          "3d": three_d, "ms": ms, "ms_best": best, "evals": evals, "evals/s": evals / (ms * 1e-3),
                          "GB/s_algorithmic": evals * 8.0 / (ms * 1e-3) / 1e9, "tables": eng.n_tables,
                          "table_build_s": build_s, "errors": int(eng.errors.item())}))


def bench_pitch():
    """C4-shaped scenes (2048^2, 1e5 spots, 3-D) at pixel pitches with and without a whole number of table samples
    per pixel: 65 nm takes the box-table / TMA path, 66.39 nm (the reference's default x241 on 16 um pixels) and
    108.33 nm (x60 on 6.5 um) gather SAT corners.  SCB_GATHER_* select the measured gather variants."""
    from bench import count_spot_pixel_evals
    size, n = 2048, 100000
    for label, pixel, mag in (("65 nm", 6.5e-6, 100), ("100 nm", 6.5e-6, 65), ("50 nm", 6.5e-6, 130),
                              ("66.39 nm", 16e-6, 241), ("108.33 nm", 6.5e-6, 60)):
        yaml = """
default:
    magnification: %d
    light_source: {angle: {value: 0.0, units: radian}}
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: %.9e, units: m}, QE: 0.73, exposure_time: 0.033}
""" % (mag, size, size, pixel)
        from scopyon_b200.engine import SatStore
        SatStore.clear_shared()
        torch.cuda.empty_cache()
        configs, eng = engine_for(yaml)
        pl = configs.pixel_length
        rng = numpy.random.RandomState(1)
        data = numpy.zeros((n, 5))
        data[:, 1:3] = rng.uniform(-size * pl / 2, size * pl / 2, (n, 2))
        data[:, 0] = rng.uniform(0, 1.5e-6, n)
        data[:, 3] = numpy.arange(n)
        data[:, 4] = 1
        t0 = time.time()
        eng.ensure_all_tables()
        torch.cuda.synchronize()
        build_s = time.time() - t0
        soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).cuda()
        w = torch.full((n,), 30.0, dtype=torch.float64, device="cuda")
        out = torch.empty((size, size), dtype=torch.float32, device="cuda")
        work = eng._render_workspace(n)

        def go():
            eng._call("scb_render_expected", ctypes.byref(eng.geom), n, _native.ptr(soa[0]), _native.ptr(soa[1]),
                      _native.ptr(soa[2]), _native.ptr(w), _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type, _native.ptr(eng.inv_scale),
                      _native.ptr(eng.slot_of_key), _native.ptr(out), _native.F32, 0, _native.ptr(work), work.numel(),
                      _native.ptr(eng.errors), eng._stream())
        ms, best = timed(go, iters=6, warm=2)
        evals = count_spot_pixel_evals(data, size, pl)
        print(json.dumps({"kernel": "scb_render_expected (prepare+scan+fill+render), one frame", "pixel": label, "size": size,
                          "spots": n, "box_table": eng.box is not None, "ms": ms, "ms_best": best, "evals": evals,
                          "evals/s": evals / (ms * 1e-3), "tables": eng.n_tables, "table_build_s": build_s,
                          "sat_GB": 0 if eng.sat is None else eng.sat.numel() * 8 / 1e9,
                          "variant": {k: v for k, v in os.environ.items() if k.startswith("SCB_")},
                          "checksum": float(out.double().sum().item()), "errors": int(eng.errors.item())}))
        del eng, soa, w, out, work


def bench_gaussian():
    """BASELINE config 5: separable-Gaussian stress, 1e6 spots, 4096 x 4096."""
    from scopyon_b200.engine import SatStore
    sys.path.insert(0, ROOT)
    from bench import count_spot_pixel_evals
    size, n = 4096, 1000000
    yaml = """
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    magnification: 100
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73}
""" % (size, size)
    rng = numpy.random.RandomState(1)
    results = {}
    for tc in (False, True):
        import warnings
        config = scopyon_b200.DefaultConfiguration()
        config.update(yaml)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
        eng = DeviceEngine(configs, precision="f32", gaussian_tc=tc)
        pl = configs.pixel_length
        data = numpy.zeros((n, 5))
        data[:, 1:3] = numpy.random.RandomState(1).uniform(-size * pl / 2, size * pl / 2, (n, 2))
        data[:, 3] = numpy.arange(n)
        data[:, 4] = 1
        snap = [(0.033, data)]
        out = torch.empty((size, size), dtype=torch.float32, device="cuda")
        eng.render_expected(snap, out=out)                      # builds tables, uploads
        torch.cuda.synchronize()
        # time the device part only: re-issue the render on resident spots
        soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).cuda()
        w = torch.full((n,), 30.0, dtype=torch.float64, device="cuda")
        if tc:
            need = eng.lib.scb_gaussian_tc_workspace_bytes(ctypes.byref(eng.geom), n)
            work = torch.empty(need + 256, dtype=torch.uint8, device="cuda")

            def go():
                eng._call("scb_render_gaussian_tc", ctypes.byref(eng.geom), n, _native.ptr(soa[1]), _native.ptr(soa[2]),
                          _native.ptr(w), _native.ptr(eng.gaussian_prefix), _native.ptr(out), _native.F32, 0,
                          _native.ptr(work), work.numel(), _native.ptr(eng.errors), eng._stream())
        else:
            def go():
                eng._render_sat(soa, w, n, out, eng._stream())
        ms, best = timed(go, iters=5, warm=2)
        results[tc] = out.double().cpu().numpy()
        evals = count_spot_pixel_evals(data, size, pl)
        print(json.dumps({"kernel": "scb_render_gaussian_tc" if tc else "scb_render_expected (SAT, Gaussian table)",
                          "config": "C5: 1e6 Gaussian spots, 4096^2", "ms": ms, "ms_best": best, "evals": evals,
                          "evals/s": evals / (ms * 1e-3), "frames/s (render only)": 1e3 / ms}))
    diff = abs(results[True] - results[False]).max() / results[False].max()
    print(json.dumps({"check": "tensor-core vs exact SAT image, max|diff|/max", "value": diff,
                      "sum_sat": float(results[False].sum()), "sum_tc": float(results[True].sum()),
                      "identical": bool((results[True] == results[False]).all())}))


if __name__ == "__main__":
    what = sys.argv[1:] or ["detector", "diffuse", "render"]
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": PEAK}))
    if "detector" in what:
        bench_detector()
    if "detector_block" in what:
        bench_detector_block()
    if "diffuse" in what:
        bench_diffuse()
    if "render" in what:
        bench_render()
    if "gaussian" in what:
        bench_gaussian()
    if "pitch" in what:
        bench_pitch()
