#!/usr/bin/env python
"""Benchmark of the image-formation hot path (BASELINE.json metric).

Workload (config.workload = "C4"): EPI 3-D movie, 1e5 molecules diffusing in
[-L/2, L/2]^2 x [0, 1.5 um], 2048 x 2048 sCMOS (CMOS read-noise table, QE 0.73, x100,
16-bit ADC, offset 100, full well 30 000, column FPN 2 counts), photobleaching on,
one snapshot per 33 ms frame.

One "step" = one block of --frames-per-step frames: emission/bleaching + Brownian steps
-> strip binning -> PSF render -> detector/ADC, everything resident in HBM.  Frames are
processed sixteen per launch (movie_frames, spot_prepare, spot_edges, tile_scan, strip_fill,
render_strips, detector_fast, detector_slow once per sixteen frames).  With N GPUs
the movie is partitioned by frame blocks (weak scaling: every rank renders the same
number of frames per step; a rank first replays the trajectory prefix of the frames
before its block, reported as replay_ms, outside the timed steps).

  python bench.py [--gpus N] [--steps K] [--warmup W]        our arm
  python bench.py --impl reference ...                       CPU arm (oracle port of the
                                                             reference algorithm, all host threads)
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C4_YAML = """
default:
    magnification: 100
    light_source: {angle: {value: 0.0, units: radian}}
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
    effects: {photo_bleaching: {switch: true, half_life: {value: %g, units: s}}}
"""
# Photobleaching stays switched on (budgets are drawn and depleted every frame), but with a
# half-life long against the half minute of movie a benchmark run covers: the metric is quoted for
# 1e5 SPOTS per frame, and at the default 2.5 s a third of the molecules would be dark -- and
# skipped, here as in the reference (_epifm.py:217-218) -- before the timed region ends.
BENCH_HALF_LIFE = 2500.0
SEED = 123
D_COEFF = 1e-13
DEPTH_MAX = 1.5e-6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--molecules", type=int, default=100000)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--e2e-frames", type=int, default=48)
    ap.add_argument("--cpu-sample-spots", type=int, default=384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_config(size):
    import scopyon_b200
    config = scopyon_b200.DefaultConfiguration()
    config.update(C4_YAML % (size, size, BENCH_HALF_LIFE))
    return config


def box(size):
    pl = 6.5e-6 / 100
    half = size * pl * 0.5
    return [-half, -half, 0.0], [half, half, DEPTH_MAX]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every few
    milliseconds from a thread (the timed region lasts tens of milliseconds, too short for
    `nvidia-smi -lms`); `nvidia-smi` is the fallback when NVML cannot be loaded."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self.proc = None
        self.thread = None
        self.running = False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; resolve it through the PCI bus id
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id
            dom = torch.cuda.get_device_properties(self.index).pci_domain_id
            dev = torch.cuda.get_device_properties(self.index).pci_device_id
            handle = pynvml.nvmlDeviceGetHandleByPciBusId("{:08x}:{:02x}:{:02x}.0".format(dom, bus, dev).encode())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, handle)
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:      # noqa: BLE001 -- any NVML problem: fall back to the CLI
            self.nvml = None
        try:
            self.rows = []
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        pynvml, handle = self.nvml
        while self.running:
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                try:
                    reasons = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
                except AttributeError:
                    reasons = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                self.samples.append((sm, reasons))
            except Exception:      # noqa: BLE001
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
            pynvml, _ = self.nvml
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = sorted(n for n in names if any(r & bits[n] for _, r in self.samples))
            sm = [v for v, _ in self.samples]
            return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": self.sm_max,
                    "samples": len(sm), "reasons": reasons, "source": "nvml, 2 ms period over the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for row in self.rows:
            cells = [c.strip() for c in row.split(",")]
            if len(cells) < 7:
                continue
            try:
                sm.append(float(cells[0]))
                smax.append(float(cells[1]))
            except ValueError:
                continue
            for name, cell in zip(names, cells[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


# --------------------------------------------------------------------------- work counting
def count_spot_pixel_evals(data, size, pl):
    """Spot-pixel evals of one frame = sum over spots of (#rows x #cols) touched, with the
    reference's footprint bounds (_epifm.py:233-235); plain numpy, used for the roofline."""
    sw = 1e-9 * 1998
    total = 0
    for col in (1, 2):
        o = size * pl * 0.5 + data[:, col] - sw * 0.5
        lo = numpy.maximum(numpy.floor(o / pl), 0)
        hi = numpy.minimum(numpy.ceil((o + sw) / pl), size)
        n = numpy.maximum(hi - lo, 0)
        total = n if col == 1 else total * n
    return float(total.sum())


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(args, n_threads):
    """Bounded sample of the workload on the host cores with the oracle port of the
    reference algorithm: per-spot slice sums over the 1999^2 PSF table (one table per
    integer-nm depth key, built on demand like the reference's cache) + per-pixel
    Poisson / categorical / ADC detector loops.  Returns (frames_per_s, detail dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import epifm_oracle as orc
    import scopyon_b200  # config layer only (host-side YAML)
    from scopyon_b200 import _epifm

    c_oracle.build()
    import warnings
    config = make_config(args.size)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    params = configs.as_oracle_params()
    geom = c_oracle.geometry(params)
    rng = numpy.random.RandomState(SEED)
    lower, upper = box(args.size)
    n = args.cpu_sample_spots
    pts = numpy.stack([rng.uniform(lower[i], upper[i], n) for i in range(3)], axis=1)   # x, y, depth
    t0 = time.perf_counter()
    # Brownian step of the sample (sampling.py:116-118)
    c_oracle.move_points(pts.copy(), numpy.sqrt(2 * D_COEFF * 0.033) * numpy.ones(3), 1)
    t_move = time.perf_counter() - t0

    keys = _epifm.depth_keys_of(pts[:, 2], params["depth_cutoff"], geom.n_depth_keys)
    n_emit = numpy.array([orc.emitted(params, d, 0.033) for d in pts[:, 2]])
    weight = numpy.array([orc.spot_weight(params, e, 1.0) for e in n_emit])
    t_table = t_render = 0.0
    expected = numpy.zeros((args.size, args.size))
    slot = numpy.full(geom.n_depth_keys + 1, -1, dtype=numpy.int32)
    uniq = numpy.unique(keys)
    batch = 8                      # tables resident at once (8 x 32 MB); the reference caches them all
    for b0 in range(0, len(uniq), batch):
        group = uniq[b0: b0 + batch]
        t0 = time.perf_counter()
        tables = numpy.stack([c_oracle.table_from_radial(orc.radial_profile(
            params, key * 1e-9 if key < geom.n_depth_keys else params["depth_cutoff"])) for key in group])
        t_table += time.perf_counter() - t0
        slot[:] = -1
        slot[group] = numpy.arange(len(group))
        sel = numpy.isin(keys, group)
        t0 = time.perf_counter()
        expected += c_oracle.render_bruteforce(geom, pts[sel, 2], pts[sel, 0], pts[sel, 1], weight[sel],
                                               tables, slot, n_threads=n_threads)
        t_render += time.perf_counter() - t0
    rn = _epifm.catalog_tables()["cmos_readout"]
    t0 = time.perf_counter()
    c_oracle.detector_frame(expected, params["QE"], params["background_mean"], True, rn["electrons"], rn["weight"],
                            0.0, params["adc_fullwell"], params["adc_offset"], params["adc_bit"], 7, n_threads=n_threads)
    t_det = time.perf_counter() - t0
    scale = args.molecules / float(n)
    # PSF tables are a one-off per depth key in the reference (cache); a long movie touches all 1002
    frame_s = (t_render + t_move) * scale + t_det
    detail = dict(sample_spots=n, render_s=t_render, table_build_s=t_table, n_tables=int(len(numpy.unique(keys))),
                  detector_s=t_det, move_s=t_move, frame_s_extrapolated=frame_s)
    return 1.0 / frame_s, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 0) and 1):
        cpu_sample(argparse.Namespace(**{**vars(args), "cpu_sample_spots": 32}), n_threads)
    values, detail = [], None
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        v, detail = cpu_sample(args, n_threads)
        values.append(v)
    value = float(numpy.mean(values))
    sample = ("{} of {} spots rendered by slice sums over per-depth 1999^2 tables + full {}^2 detector loop, "
              "spot cost scaled linearly; PSF-table build ({:.2f} s for {} keys) excluded as one-off").format(
                  detail["sample_spots"], args.molecules, args.size, detail["table_build_s"], detail["n_tables"])
    line = {
        "impl": "reference", "metric": "frames/sec (2048^2 sCMOS, 1e5 spots)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "C4: EPI 3-D diffusion, {} molecules, {}x{} sCMOS (CMOS table noise, column FPN), "
                        "photobleaching on (half-life {:g} s: >= 98 % of the molecules emit in every timed frame), "
                        "1 snapshot/frame".format(args.molecules, args.size, args.size, BENCH_HALF_LIFE),
            "frames_per_step": args.frames_per_step, "molecules": args.molecules,
            "image_size": [args.size, args.size], "parallelism": "frame-blocks x{}".format(args.gpus),
            "cache": "per-frame working set (35 GB of PSF box tables read at random, ~0.8 GB per frame) "
                     "exceeds the 126 MB L2"}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from scopyon_b200 import _native
    from scopyon_b200.movie import DeviceMovie

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    lib = _native.load()

    F, K, W = args.frames_per_step, args.steps, args.warmup
    config = make_config(args.size)
    lower, upper = box(args.size)
    t0 = time.perf_counter()
    movie = DeviceMovie(config, args.molecules, lower, upper, D_COEFF, SEED, device=device, precision="f32")
    if os.environ.get("SCB_FRAMES_PER_LAUNCH"):
        movie.frames_per_launch = int(os.environ["SCB_FRAMES_PER_LAUNCH"])
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    frames_per_rank = (W + K) * F
    first = rank * frames_per_rank
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    movie.reset(first_frame=first)
    ev1.record()
    torch.cuda.synchronize()
    replay_ms = ev0.elapsed_time(ev1)

    block = torch.empty((F, args.size, args.size), dtype=torch.float32, device=device)
    for _ in range(W):
        movie.render_block(block)
    torch.cuda.synchronize()
    evals = count_spot_pixel_evals(movie.positions(), args.size, 6.5e-6 / 100)
    emitting_start = float((movie.weight > 0).double().mean().item())    # molecules not bleached yet

    sampler = ClockSampler(local)
    sampler.start()
    _native.check(lib.scb_profile_begin(K * F), "scb_profile_begin")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    t_enqueue = time.perf_counter()
    for _ in range(K):
        movie.render_block(block)
    t_enqueue = time.perf_counter() - t_enqueue        # host time to enqueue the timed frames
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = start.elapsed_time(stop)
    render_ms, render_launches = ctypes.c_double(0), ctypes.c_int64(0)
    _native.check(lib.scb_profile_end(ctypes.byref(render_ms), ctypes.byref(render_launches)), "scb_profile_end")
    clocks = sampler.stop()
    emitting_end = float((movie.weight > 0).double().mean().item())
    evals *= 0.5 * (emitting_start + emitting_end)       # dark molecules are skipped (_epifm.py:217-218)
    n_err = int(movie.engine.errors.item())
    checksum = float(block[-1].double().mean().item())

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * K * F / (elapsed_ms * 1e-3)

    # ---- end to end through the public API: host positions in, float64 frames out
    e2e = run_e2e(args, config, movie, world, device)

    if rank == 0:
        peaks = {}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peaks = json.load(open(peaks_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_ms = render_ms.value / max(1, render_launches.value)
        frames_per_launch = K * F / max(1, render_launches.value)      # render_block renders several frames per launch
        evals *= frames_per_launch
        # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture
        traffic = None
        traffic_path = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(traffic_path) and args.molecules == 100000 and args.size == 2048:
            for name, entry in json.load(open(traffic_path)).items():
                if name.startswith("render_strips_kernel<float"):
                    traffic = entry.get("dram_bytes_per_launch")
        # algorithmic bytes: one box-table value per spot-pixel eval -- 4 bytes with
        # the fp32 tables that fp32 frames use, 8 with fp64 tables
        bytes_per_eval = 4.0 if movie.engine.box is not None and movie.engine.box.dtype == torch.float32 else 8.0
        achieved = evals * bytes_per_eval / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "frames/sec (2048^2 sCMOS, 1e5 spots)", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 box tables and frames, 32-bit fixed-point accumulation per strip (edge arithmetic f64)",
            "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e,
            # per block of sixteen frames: movie_frames, spot_prepare, spot_edges, tile_scan, strip_fill,
            # render_strips, detector_fast, detector_slow
            "gpu_launches": int(8 * render_launches.value),
            "roofline": {
                "kernel": "render_strips_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": evals * bytes_per_eval, "bytes_per_spot_pixel_eval": bytes_per_eval,
                "spot_pixel_evals_per_launch": evals,
                "spot_pixel_evals_per_s": evals / (per_launch_ms * 1e-3), "ms_per_launch": per_launch_ms,
                "frames_per_launch": frames_per_launch,
                "share_of_step": render_ms.value / elapsed_ms,
            },
            "emitting_fraction": [emitting_start, emitting_end],
            "host_enqueue_ms_per_frame": t_enqueue * 1e3 / (K * F),
            "replay_ms": replay_ms, "setup_s": setup_s, "frame_checksum_mean_adc": checksum, "table_errors": n_err,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, detail = cpu_sample(args, os.cpu_count() or 1)
            line["cpu_baseline"] = {
                "value": v, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                "sample": "{} of {} spots by per-pixel slice sums over per-depth 1999^2 tables + full-frame detector "
                          "loop, spot cost scaled linearly (render {:.2f} s, detector {:.2f} s; table build {:.2f} s "
                          "for {} keys excluded as one-off)".format(
                              detail["sample_spots"], args.molecules, detail["render_s"], detail["detector_s"],
                              detail["table_build_s"], detail["n_tables"])}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, config, movie, world, device):
    """The same metric through scopyon_b200.generate_images: per frame the (N, 5) float64
    positions go host -> device from pinned memory and the frame comes back (float32 on the
    wire, float64 in the caller's hands)."""
    import torch
    import torch.distributed as dist
    import scopyon_b200
    n_frames = args.e2e_frames
    # host trajectory: positions of consecutive frames taken from the device movie
    inputs = []
    block = torch.empty((1, args.size, args.size), dtype=torch.float32, device=device)
    for k in range(n_frames + 4):
        inputs.append((k * 0.033, movie.positions()[:, [1, 2, 0, 3, 4]]))   # (x, y, z, id, p_state) rows
        movie.render_block(block)
    rng = numpy.random.RandomState(SEED + 1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = scopyon_b200.create_simulator(config, rng=rng)
        gen = sim.generate_images(inputs, num_frames=n_frames + 4)
        first = next(gen)                      # warm-up frames: build/attach PSF tables, allocate buffers
        for _ in range(3):
            first = next(gen)
        del first
        from scopyon_b200.engine import DeviceEngine
        DeviceEngine.trace = {}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        total = 0.0
        for img in gen:
            total += float(img.as_array()[0, 0])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    breakdown = {k: v / n_frames for k, v in (DeviceEngine.trace or {}).items()}
    DeviceEngine.trace = None
    return {"value": world * n_frames / dt, "unit": "frames/s", "host_ms_per_frame": breakdown,
            "h2d_bytes_per_step": int(args.molecules * (4 * 8 + 4 + 8)),
            "d2h_bytes_per_step": int(args.size * args.size * 4),
            "d2h": "float32 frame into pinned staging memory, widened to the float64 array the API returns by host threads (scb_host_widen_*)",
            "frames_timed": n_frames, "per": "frame (one generate_images iteration)"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
