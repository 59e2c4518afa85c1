#!/usr/bin/env python
"""Benchmark of the image-formation hot path (BASELINE.json metric).

Workload C4 (configs[3], the one the metric is quoted on): EPI 3-D movie, 1e5 molecules diffusing in
[-L/2, L/2]^2 x [0, 1.5 um], 2048 x 2048 sCMOS (CMOS read-noise table, QE 0.73, x100, 16-bit ADC,
offset 100, full well 30 000, column FPN 2 counts), photobleaching on, one snapshot per 33 ms frame
(SURVEY.md section 8(d)).

  python bench.py [--gpus N] [--steps K] [--warmup W]        weak scaling (the driver's call): a step is
        a block of --frames-per-step frames per rank, resident in HBM; the JSON line also carries the
        rate of the data plane (every finished block streamed to page-locked host memory as float32 /
        uint16 / uint8, `export`), the end-to-end rate through generate_images (`e2e`), the host bounds
        they sit under (`host`), a parity check of a bench frame by the oracle (`parity_ok`, in a
        subprocess) and, with N > 1, a bit-for-bit check of the NCCL frame gather (`gather_ok`)
  python bench.py --scaling strong [--movie-frames 10000]    ONE fixed movie partitioned by frame blocks
        over the N ranks: trajectory replay of the prefix + rendering + export of every frame in the
        timed region
  python bench.py --workload C5                              1e6 Gaussian spots on 4096^2 (configs[4]),
        box-table path and tcgen05 path
  python bench.py --impl reference ...                       CPU arm (oracle port of the reference
        algorithm on all host threads; the live reference's own timings, taken in the build
        container, ride along from profiles/reference_live_r2.json)
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
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
C5_YAML = """
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    magnification: 100
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
"""
# Photobleaching stays switched on (budgets are drawn and depleted every frame).  The headline line uses
# a half-life long against the movie a run covers: the metric is quoted for 1e5 SPOTS per frame, and at
# the default 2.5 s a third of the molecules would be dark -- and skipped, here as in the reference
# (_epifm.py:217-218) -- before the timed region ends.  The 2.5 s of SURVEY.md section 8(d) is measured
# too and reported beside it (`spec_half_life`).
BENCH_HALF_LIFE = 2500.0
SPEC_HALF_LIFE = 2.5
SEED = 123
D_COEFF = 1e-13
DEPTH_MAX = 1.5e-6
METRIC = "frames/sec (2048^2 sCMOS, 1e5 spots)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-check"])
    ap.add_argument("--workload", default="C4", choices=["C4", "C5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--movie-frames", type=int, default=10000, help="strong scaling: frames of the one movie")
    ap.add_argument("--export", default="f32", choices=["f32", "u16", "u8"],
                    help="strong scaling: format the headline value is exported in (all three are timed)")
    ap.add_argument("--export-frames", type=int, default=384, help="weak scaling: frames per rank per export format")
    ap.add_argument("--half-life", type=float, default=BENCH_HALF_LIFE)
    ap.add_argument("--frames-per-step", type=int, default=128)
    ap.add_argument("--molecules", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--e2e-frames", type=int, default=96)
    ap.add_argument("--cpu-sample-spots", type=int, default=384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true",
                    help="kernel work only: skip export, end-to-end, host bounds, parity and CPU legs (variant sweeps)")
    ap.add_argument("--payload", default=None, help="(internal) parity payload of the cpu-check leg")
    args = ap.parse_args()
    if args.molecules is None:
        args.molecules = 1000000 if args.workload == "C5" else 100000
    if args.size is None:
        args.size = 4096 if args.workload == "C5" else 2048
    return args


def make_config(size, half_life=BENCH_HALF_LIFE):
    import scopyon_b200
    config = scopyon_b200.DefaultConfiguration()
    config.update(C4_YAML % (size, size, half_life))
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
def count_spot_pixel_evals(data, size, pl, sw=1e-9 * 1998):
    """Spot-pixel evals of one frame = sum over spots of (#rows x #cols) touched, with the
    reference's footprint bounds (_epifm.py:233-235); plain numpy, used for the roofline."""
    total = 0
    for col in (1, 2):
        o = size * pl * 0.5 + data[:, col] - sw * 0.5
        lo = numpy.maximum(numpy.floor(o / pl), 0)
        hi = numpy.minimum(numpy.ceil((o + sw) / pl), size)
        n = numpy.maximum(hi - lo, 0)
        total = n if col == 1 else total * n
    return float(total.sum())


# --------------------------------------------------------------------------- CPU arm
def cpu_setup(args, n_threads):
    """State of the CPU arm: the oracle port of the reference algorithm (per-spot slice sums over the 1999^2 PSF
    table of the spot's integer-nm depth key + per-pixel Poisson / categorical / ADC detector loops) on a BOUNDED
    SAMPLE of the workload -- `--cpu-sample-spots` of the molecules, placed like the workload's in x and y and on 16
    depth keys spread over its depth range (the cost of a spot in the reference does not depend on its depth; 16
    tables of 32 MB stay resident across steps the way the reference's PSF cache keeps them)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import epifm_oracle as orc
    import scopyon_b200  # noqa: F401 -- config layer only (host-side YAML)
    from scopyon_b200 import _epifm

    c_oracle.build()
    import warnings
    config = make_config(args.size, args.half_life)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    params = configs.as_oracle_params()
    geom = c_oracle.geometry(params)
    rng = numpy.random.RandomState(SEED)
    lower, upper = box(args.size)
    n = args.cpu_sample_spots
    pts = numpy.stack([rng.uniform(lower[i], upper[i], n) for i in range(3)], axis=1)   # x, y, depth
    levels = numpy.linspace(lower[2], upper[2], 16, endpoint=False) + 0.5e-9
    pts[:, 2] = levels[numpy.arange(n) % 16]
    keys = _epifm.depth_keys_of(pts[:, 2], params["depth_cutoff"], geom.n_depth_keys)
    n_emit = numpy.array([orc.emitted(params, d, 0.033) for d in pts[:, 2]])
    weight = numpy.array([orc.spot_weight(params, e, 1.0) for e in n_emit])
    uniq = numpy.unique(keys)
    t0 = time.perf_counter()
    tables = numpy.stack([c_oracle.table_from_radial(orc.radial_profile(
        params, key * 1e-9 if key < geom.n_depth_keys else params["depth_cutoff"])) for key in uniq])
    t_table = time.perf_counter() - t0
    slot = numpy.full(geom.n_depth_keys + 1, -1, dtype=numpy.int32)
    slot[uniq] = numpy.arange(len(uniq))
    rn = _epifm.catalog_tables()["cmos_readout"]
    return dict(c_oracle=c_oracle, params=params, geom=geom, pts=pts, weight=weight, tables=tables, slot=slot, rn=rn,
                table_build_s=t_table, n_tables=int(len(uniq)), n=n, size=args.size, molecules=args.molecules)


def cpu_step(state, n_threads):
    """One step of the CPU arm: Brownian step, overlay and detector pass of the sample (one frame).
    Returns (move_s, render_s, detector_s)."""
    c_oracle, params, pts = state["c_oracle"], state["params"], state["pts"]
    t0 = time.perf_counter()
    c_oracle.move_points(pts.copy(), numpy.sqrt(2 * D_COEFF * 0.033) * numpy.ones(3), 1)     # sampling.py:116-118
    t_move = time.perf_counter() - t0
    t0 = time.perf_counter()
    expected = c_oracle.render_bruteforce(state["geom"], pts[:, 2], pts[:, 0], pts[:, 1], state["weight"],
                                          state["tables"], state["slot"], n_threads=n_threads)
    t_render = time.perf_counter() - t0
    rn = state["rn"]
    t0 = time.perf_counter()
    c_oracle.detector_frame(expected, params["QE"], params["background_mean"], True, rn["electrons"], rn["weight"],
                            0.0, params["adc_fullwell"], params["adc_offset"], params["adc_bit"], 7, n_threads=n_threads)
    t_det = time.perf_counter() - t0
    return t_move, t_render, t_det


def cpu_extrapolate(state, t_move, t_render, t_det):
    """Seconds per FULL frame from the sample's times: spot costs scale with the number of spots, the detector pass
    already covers the whole frame; PSF tables are a one-off per depth key in the reference (cache) and are left out."""
    scale = state["molecules"] / float(state["n"])
    return (t_render + t_move) * scale + t_det


def cpu_sample(args, n_threads):
    """One step of the CPU arm (the GPU arm's `cpu_baseline` leg).  Returns (frames_per_s, detail dict)."""
    state = cpu_setup(args, n_threads)
    t_move, t_render, t_det = cpu_step(state, n_threads)
    frame_s = cpu_extrapolate(state, t_move, t_render, t_det)
    detail = dict(sample_spots=state["n"], render_s=t_render, table_build_s=state["table_build_s"],
                  n_tables=state["n_tables"], detector_s=t_det, move_s=t_move, frame_s_extrapolated=frame_s)
    return 1.0 / frame_s, detail


def reference_live():
    """Timings of the UNMODIFIED reference taken in the build container
    by tools/time_reference.py -- /root/reference does not travel to the GPU box, so they ride along as a
    committed record, labelled with the machine they were measured on."""
    path = os.path.join(ROOT, "profiles", "reference_live_r2.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def parity_check(payload_path):
    """Oracle check of a bench frame through 40 of its spots: (frame - frame without them), rendered by
    the GPU arm, against the oracle's expected image of those 40 (linearity; the same check as
    tests/test_gpu_fullsize.py at full size)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import epifm_oracle as orc
    with numpy.load(payload_path, allow_pickle=True) as z:
        params = z["params"].item()
        picked, difference, alone, peak = z["picked"], z["difference"], z["alone"], float(z["peak"])
    want, _ = orc.expected_frame([(0.0, picked)], params, exposure_time=0.033)
    err_alone = float(abs(alone - want).max() / want.max())
    err_diff = float(abs(difference - want).max() / peak)
    same_footprint = bool(((alone > 0) == (want > 0)).all())
    # tolerances of tests/test_gpu_fullsize.py (fp32 tables and accumulators)
    ok = err_alone < 5e-7 and err_diff < 1e-6 and same_footprint
    return {"parity_ok": bool(ok), "spots": int(len(picked)), "max_err_subset_over_max": err_alone,
            "max_err_frame_difference_over_frame_max": err_diff, "identical_footprints": same_footprint,
            "tolerance": "5e-7 of the subset's maximum / 1e-6 of the frame's maximum (fp32 mode)"}


def run_cpu_check(args):
    """Subprocess leg of the GPU arm: everything that touches oracle/ (the CPU baseline and the parity
    check) runs here, so the GPU arm's process maps only the product's library."""
    out = {}
    if args.payload:
        out["parity"] = parity_check(args.payload)
    if not args.no_cpu_baseline:
        v, detail = cpu_sample(args, os.cpu_count() or 1)
        out["cpu_baseline"] = {
            "value": v, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "{} of {} spots by per-pixel slice sums over per-depth 1999^2 tables + full-frame detector "
                      "loop, spot cost scaled linearly (render {:.2f} s, detector {:.2f} s; table build {:.2f} s "
                      "for {} keys excluded as one-off)".format(
                          detail["sample_spots"], args.molecules, detail["render_s"], detail["detector_s"],
                          detail["table_build_s"], detail["n_tables"])}
    print(json.dumps(out))


def cpu_check_subprocess(args, payload_path):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "cpu-check", "--size", str(args.size),
           "--molecules", str(args.molecules), "--cpu-sample-spots", str(args.cpu_sample_spots),
           "--half-life", str(args.half_life)]
    if payload_path:
        cmd += ["--payload", payload_path]
    if args.no_cpu_baseline:
        cmd += ["--no-cpu-baseline"]
    env = dict(os.environ)
    for key in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(key, None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    if res.returncode != 0:
        return {"error": (res.stderr or res.stdout)[-400:]}
    return json.loads(res.stdout.strip().splitlines()[-1])


def run_reference(args):
    """The reference arm: W untimed and EXACTLY K timed steps of the CPU port on the bounded sample (`cpu_setup`), all
    host threads.  `ms_per_step` is the measured wall time of one sample step; `value` is the metric extrapolated from
    the steps' times to the full workload (`cpu_extrapolate`), and says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    state = cpu_setup(args, n_threads)
    for _ in range(max(args.warmup, 0)):
        cpu_step(state, n_threads)
    times, walls = [], []
    for _ in range(max(1, args.steps)):
        w0 = time.perf_counter()
        times.append(cpu_step(state, n_threads))
        walls.append(time.perf_counter() - w0)
    t_move, t_render, t_det = (float(numpy.mean([t[i] for t in times])) for i in range(3))
    value = 1.0 / cpu_extrapolate(state, t_move, t_render, t_det)
    sample = ("one step = one frame of {} of the {} spots (16 depth keys) rendered by slice sums over per-depth 1999^2 "
              "tables + Brownian step + the full {}^2 detector loop; value = 1 / (spot seconds x {:.1f} + detector "
              "seconds); PSF-table build ({:.2f} s for {} keys) is a one-off and not in the steps").format(
                  state["n"], args.molecules, args.size, args.molecules / float(state["n"]), state["table_build_s"],
                  state["n_tables"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(args.warmup, 0),
        "ms_per_step": 1e3 * float(numpy.mean(walls)),
        "extrapolated_ms_per_frame": 1e3 / value,
        "step_seconds": {"brownian": t_move, "overlay": t_render, "detector": t_det},
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": n_threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    live = reference_live()
    if live is not None:
        line["reference_live"] = live
    print(json.dumps(line))


def workload_config(args):
    emitting = ">= 98 % of the molecules emit in every timed frame" if args.half_life >= 100 else \
        "molecules bleach during the run: see emitting_fraction"
    cfg = {"workload": "C4: EPI 3-D diffusion, {} molecules, {}x{} sCMOS (CMOS table noise, column FPN), "
                       "photobleaching on (half-life {:g} s: {}), 1 snapshot/frame".format(
                           args.molecules, args.size, args.size, args.half_life, emitting),
           "frames_per_step": args.frames_per_step, "molecules": args.molecules,
           "image_size": [args.size, args.size], "parallelism": "frame-blocks x{}".format(args.gpus),
           "cache": "per-frame working set (17 GB of fp32 PSF box tables read at random, ~0.4 GB per frame) "
                    "exceeds the 126 MB L2"}
    if args.scaling == "strong":
        cfg["movie_frames"] = args.movie_frames
        cfg["frames_per_step"] = None
    return cfg


# --------------------------------------------------------------------------- GPU arm: helpers
def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local, torch.device("cuda", local)


def max_over_ranks(value, world, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, world, device):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))
    return {}


def host_bounds(lib, world, device, frame_bytes):
    """The bounds the host side of a box puts on frames leaving its GPUs: page-locked device->host
    bandwidth with every rank copying at once, and the memory bandwidth of the host cores (copy and
    float32 -> float64 widening patterns, scb_host_bandwidth) with this rank's share of the threads."""
    import torch
    n_bytes = 256 << 20
    src = torch.empty(n_bytes, dtype=torch.uint8, device=device)
    dst = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 6
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), world, device)
    d2h = world * reps * n_bytes / (ms * 1e-3) / 1e9
    del src, dst
    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    out = {"pcie_d2h_gbs_all_ranks": d2h, "frames_per_s_at_pcie_f32": d2h * 1e9 / frame_bytes,
           "threads_per_rank": threads, "host_cores": os.cpu_count()}
    for mode, name in ((0, "host_copy_gbs_all_ranks"), (1, "host_widen_gbs_all_ranks")):
        v = ctypes.c_double(0.0)
        barrier(world)
        rc = lib.scb_host_bandwidth(mode, 64 << 20, threads, 3, ctypes.byref(v))
        out[name] = sum_over_ranks(v.value / 1e9 if rc == 0 else 0.0, world, device)
    # float64 frames: 4 B/px DMA write + 4 B/px read + 8 B/px write on the host per frame
    out["frames_per_s_at_host_widen_f64"] = out["host_widen_gbs_all_ranks"] * 1e9 / (3 * frame_bytes)
    return out


def export_rate(movie, fmt, frames, world, device):
    """Frames per second of the data plane: `frames` frames per rank rendered AND streamed to page-locked
    host memory (DeviceMovie.stream_frames), device-timed from the first launch to the last download."""
    import torch
    touched = [0.0]

    def sink(first, block):
        touched[0] += float(block[0, 0, 0])      # the consumer reads from every delivered block

    movie.stream_frames(2 * movie.export_block_frames, fmt=fmt, sink=sink)     # buffers, pinned ring
    barrier(world)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    t0 = time.perf_counter()
    n = movie.stream_frames(frames, fmt=fmt, sink=sink)
    wall = time.perf_counter() - t0          # stream_frames returns after the last block was delivered
    stop.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(max(start.elapsed_time(stop), wall * 1e3), world, device)
    bytes_per_px = {"f32": 4, "u16": 2, "u8": 1}[fmt]
    eng = movie.engine
    return {"value": world * n / (ms * 1e-3), "unit": "frames/s", "frames_per_rank": n,
            "d2h_bytes_per_frame": eng.n_w * eng.n_h * bytes_per_px}


def gather_check(args, world, rank, device):
    """A small movie rendered by frame blocks on the N ranks, gathered with movie.gather_frames (NCCL
    all_gather), must equal the same movie rendered by one rank, bit for bit."""
    import torch
    import torch.distributed as dist
    from scopyon_b200.movie import DeviceMovie, frame_block, gather_frames
    if world < 2:
        return None
    size, n_mol, total = 256, 3000, 5 * world + 3          # ragged blocks
    config = make_config(size, args.half_life)
    lower, upper = box(size)
    movie = DeviceMovie(config, n_mol, lower, upper, D_COEFF, SEED + 7, device=device, precision="f32")
    first, last = frame_block(total, rank, world)
    movie.reset(first_frame=first)
    mine = torch.empty((last - first, size, size), dtype=torch.float32, device=device)
    movie.render_block(mine)
    stack = gather_frames(mine, total)
    movie.reset(first_frame=0)
    whole = torch.empty((total, size, size), dtype=torch.float32, device=device)
    movie.render_block(whole)
    ok = torch.tensor([1.0 if torch.equal(stack, whole) else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item() == 1.0)


def parity_payload(movie, args):
    """Renders the oracle-subset check of the movie's current frame (expected image with all molecules,
    without 40 of them, and of the 40 alone) and stores what the oracle needs."""
    import torch
    eng = movie.engine
    data = movie.positions()
    rng = numpy.random.RandomState(11)
    pick = rng.choice(len(data), 40, replace=False)
    pick[:4] = numpy.argsort(data[:, 0])[[0, 1, -2, -1]]         # shallowest and deepest molecules
    rest = numpy.ones(len(data), dtype=bool)
    rest[pick] = False

    def render(rows):
        out = torch.empty((eng.n_w, eng.n_h), dtype=torch.float32, device=eng.device)
        img, _ = eng.render_expected([(0.033, rows)], out=out)
        torch.cuda.synchronize()
        return img.cpu().numpy().astype(numpy.float64)

    whole = render(data)
    without = render(data[rest])
    alone = render(data[pick])
    path = os.path.join(tempfile.gettempdir(), "scb_bench_parity_{}.npz".format(os.getpid()))
    numpy.savez(path, params=numpy.array(movie.configs.as_oracle_params(), dtype=object), picked=data[pick],
                difference=whole - without, alone=alone, peak=whole.max())
    return path


def timed_blocks(movie, block, steps, lib, world):
    """`steps` calls of render_block, device-timed; returns (elapsed ms of this rank, render kernel ms,
    render launches, host enqueue seconds)."""
    import torch
    from scopyon_b200 import _native
    _native.check(lib.scb_profile_begin(steps * block.shape[0]), "scb_profile_begin")
    barrier(world)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    t_enqueue = time.perf_counter()
    for _ in range(steps):
        movie.render_block(block)
    t_enqueue = time.perf_counter() - t_enqueue
    stop.record()
    barrier(world)
    render_ms, render_launches = ctypes.c_double(0), ctypes.c_int64(0)
    _native.check(lib.scb_profile_end(ctypes.byref(render_ms), ctypes.byref(render_launches)), "scb_profile_end")
    return start.elapsed_time(stop), render_ms.value, int(render_launches.value), t_enqueue


def traffic_record(args):
    """DRAM bytes per launch of the render kernel from the committed `ncu --set full` capture."""
    if args.molecules != 100000 or args.size != 2048:
        return None, None, None
    for name in ("traffic_r2.json", "traffic_r1.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            for kernel, entry in json.load(open(path)).items():
                if kernel.startswith("render_strips_kernel<float"):
                    return entry.get("dram_bytes_per_launch"), "profiles/" + name, entry.get("frames_per_launch")
    return None, None, None


# --------------------------------------------------------------------------- GPU arm: weak scaling (default)
def run_weak(args):
    import torch
    import torch.distributed as dist
    from scopyon_b200 import _native
    from scopyon_b200.movie import DeviceMovie

    world, rank, local, device = setup_dist()
    lib = _native.load()
    F, K, W = args.frames_per_step, args.steps, args.warmup
    config = make_config(args.size, args.half_life)
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
    replay_ms = max_over_ranks(ev0.elapsed_time(ev1), world, device)       # the last rank replays the longest prefix

    block = torch.empty((F, args.size, args.size), dtype=torch.float32, device=device)
    for _ in range(W):
        movie.render_block(block)
    torch.cuda.synchronize()
    evals = count_spot_pixel_evals(movie.positions(), args.size, 6.5e-6 / 100)
    emitting_start = float((movie.weight > 0).double().mean().item())    # molecules not bleached yet

    sampler = ClockSampler(local)
    sampler.start()
    elapsed_ms, render_ms, render_launches, t_enqueue = timed_blocks(movie, block, K, lib, world)
    clocks = sampler.stop()
    emitting_end = float((movie.weight > 0).double().mean().item())
    evals *= 0.5 * (emitting_start + emitting_end)       # dark molecules are skipped (_epifm.py:217-218)
    n_err = int(movie.engine.errors.item())
    checksum = float(block[-1].double().mean().item())
    elapsed_ms = max_over_ranks(elapsed_ms, world, device)
    value = world * K * F / (elapsed_ms * 1e-3)

    if args.resident_only:
        if rank == 0:
            per_launch_ms = render_ms / max(1, render_launches)
            print(json.dumps({"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
                              "warmup": W, "ms_per_step": elapsed_ms / K, "render_ms_per_launch": per_launch_ms,
                              "render_share_of_step": render_ms / elapsed_ms,
                              "variant": {k: v for k, v in os.environ.items() if k.startswith("SCB_")},
                              "frame_checksum_mean_adc": checksum, "table_errors": n_err, "clocks": clocks}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- the 2.5 s half-life of SURVEY.md section 8(d), beside the headline: the first F frames of the movie
    # (4.2 s of it: two thirds of the molecules bleach on the way and are skipped from then on)
    spec = None
    if args.half_life != SPEC_HALF_LIFE:
        movie_spec = DeviceMovie(make_config(args.size, SPEC_HALF_LIFE), args.molecules, lower, upper, D_COEFF, SEED,
                                 device=device, precision="f32")
        movie_spec.frames_per_launch = movie.frames_per_launch
        movie_spec.render_block(block)                         # buffers
        movie_spec.reset(first_frame=0)
        ms, _, _, _ = timed_blocks(movie_spec, block, 1, lib, world)
        ms = max_over_ranks(ms, world, device)
        e1 = float((movie_spec.weight > 0).double().mean().item())
        spec = {"half_life_s": SPEC_HALF_LIFE, "value": world * F / (ms * 1e-3), "unit": "frames/s",
                "frames": "0 .. {} of the movie on every rank".format(F - 1), "emitting_fraction": [1.0, e1],
                "note": "bleached molecules are skipped (as in the reference): fewer spots per frame than the metric names"}
        del movie_spec

    # ---- the data plane: every finished block streamed to page-locked host memory
    export = {fmt: export_rate(movie, fmt, args.export_frames, world, device) for fmt in ("f32", "u16", "u8")}
    # ---- parity of a bench frame (oracle subset, checked in the CPU subprocess)
    payload = parity_payload(movie, args) if rank == 0 else None
    # ---- NCCL gather of a frame-block partitioned movie == single-rank movie
    gather_ok = gather_check(args, world, rank, device)
    # ---- host-side bounds
    host = host_bounds(lib, world, device, args.size * args.size * 4)
    # ---- end to end through the public API: host positions in, frames out
    e2e = run_e2e(args, config, movie, world, device)

    if rank == 0:
        peaks = load_peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_ms = render_ms / max(1, render_launches)
        frames_per_launch = K * F / max(1, render_launches)      # render_block renders several frames per launch
        evals *= frames_per_launch
        # the capture is one launch of `captured_frames` frames; the timed launches hold frames_per_launch: per launch, like `achieved`
        traffic, traffic_source, captured_frames = traffic_record(args)
        if traffic is not None and captured_frames:
            traffic = traffic * frames_per_launch / captured_frames
            traffic_source += " (%d-frame launch, scaled to %g frames per launch)" % (captured_frames, frames_per_launch)
        # algorithmic bytes: one box-table value per spot-pixel eval -- 4 bytes with
        # the fp32 tables that fp32 frames use, 8 with fp64 tables
        bytes_per_eval = 4.0 if movie.engine.box is not None and movie.engine.box.dtype == torch.float32 else 8.0
        achieved = evals * bytes_per_eval / (per_launch_ms * 1e-3) / 1e9
        for bound_key, rate_key in (("frames_per_s_at_pcie_f32", "value"), ("frames_per_s_at_host_widen_f64", "float64_value")):
            if host.get(bound_key) and e2e.get(rate_key):
                e2e["frac_of_" + bound_key] = e2e[rate_key] / host[bound_key]
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": elapsed_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 box tables and frames, 32-bit fixed-point accumulation per strip (edge arithmetic f64)",
            "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e,
            # per launch group (frames_per_launch frames) of a planned block: movie_frames, spot_bin_fused, spot_edges,
            # render_strips, tile_scan (the next block's list plan), detector_fast, detector_slow
            "gpu_launches": int(7 * render_launches),
            "roofline": {
                "kernel": "render_strips_kernel<float>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": evals * bytes_per_eval, "bytes_per_spot_pixel_eval": bytes_per_eval,
                "spot_pixel_evals_per_launch": evals,
                "spot_pixel_evals_per_s": evals / (per_launch_ms * 1e-3), "ms_per_launch": per_launch_ms,
                "frames_per_launch": frames_per_launch,
                "share_of_step": render_ms / elapsed_ms,
                # what keeps it below the HBM roofline: profiles/render_variants_r2.md
                "limiter": "co-limited: shared-memory wavefronts 83 % of peak (box load, accumulator load and store per pixel "
                           "row, the copies' writes), issue slots 82 %, and the TMA unit's request rate -- one bulk copy per unit "
                           "at 36 cycles per unit per SM, where back-to-back bulk copies alone retire one per 40-50 cycles whatever "
                           "their size (tools/probes/tma_rate_probe.cu); DRAM 57 % active (profiles/traffic_r2.json, "
                           "profiles/ncu_block_r2_summary.csv)",
            },
            "export": export, "host": host, "gather_ok": gather_ok, "spec_half_life": spec,
            "emitting_fraction": [emitting_start, emitting_end],
            "host_enqueue_ms_per_frame": t_enqueue * 1e3 / (K * F),
            "replay_ms": replay_ms, "setup_s": setup_s, "frame_checksum_mean_adc": checksum, "table_errors": n_err,
        }
        checked = cpu_check_subprocess(argparse.Namespace(**{**vars(args), "no_cpu_baseline": args.no_cpu_baseline or world > 1}),
                                       payload)
        if "parity" in checked:
            line["parity_ok"] = checked["parity"]["parity_ok"]
            line["parity"] = checked["parity"]
        if "cpu_baseline" in checked:
            line["cpu_baseline"] = checked["cpu_baseline"]
        if "error" in checked:
            line["cpu_check_error"] = checked["error"]
        if payload and os.path.exists(payload):
            os.remove(payload)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, config, movie, world, device):
    """The same metric through scopyon_b200.generate_images: per frame the (N, 5) float64 positions go
    host -> device from pinned memory and the frame comes back into host memory.  `value`: the frame is
    in the caller's hands as the Image generate_images yields -- a float32 payload in page-locked memory
    (exact; Image.as_array() widens it to the reference's float64 array on demand) -- and one pixel is
    read from it.  `float64_value`: the caller asks every frame for its float64 array."""
    import torch
    import scopyon_b200
    from scopyon_b200.engine import DeviceEngine
    n_frames = args.e2e_frames
    # host trajectory: positions of consecutive frames taken from the device movie
    inputs = []
    block = torch.empty((1, args.size, args.size), dtype=torch.float32, device=device)
    # warm-up frames: PSF tables, buffers, and -- frames travel in blocks of engine.BLOCK_FRAMES with the next block
    # enqueued while this one is handed out -- the page-locked payloads of three blocks
    from scopyon_b200 import engine as engine_module
    n_warm = max(4, 3 * engine_module.BLOCK_FRAMES)
    for k in range(n_frames + n_warm):
        inputs.append((k * 0.033, movie.positions()[:, [1, 2, 0, 3, 4]]))   # (x, y, z, id, p_state) rows
        movie.render_block(block)
    import warnings
    from scopyon_b200.sampling import DevicePoints
    ids = numpy.arange(args.molecules, dtype=numpy.int64)
    resident = [(t, DevicePoints(torch.from_numpy(p).to(device), ids)) for t, p in inputs]
    out = {}
    for key, as_dtype, source in (("value", numpy.float32, inputs), ("float64_value", None, inputs),
                                  ("device_inputs_value", numpy.float32, resident)):
        rng = numpy.random.RandomState(SEED + 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = scopyon_b200.create_simulator(config, rng=rng)
            gen = sim.generate_images(source, num_frames=n_frames + n_warm)
            for _ in range(n_warm):                # warm-up frames: build/attach PSF tables, allocate buffers --
                first = next(gen)                  # consumed like the timed ones, so one-off allocations of the
                first.as_array(as_dtype)           # route this consumer takes happen here
            del first
            DeviceEngine.trace = {}
            barrier(world)
            t0 = time.perf_counter()
            total = 0.0
            for img in gen:
                total += float(img.as_array(as_dtype)[0, 0])
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, world, device)
        out[key] = world * n_frames / dt
        if key == "value":
            out["host_ms_per_frame"] = {k: v / n_frames for k, v in (DeviceEngine.trace or {}).items()}
        DeviceEngine.trace = None
    out.update({
        "unit": "frames/s",
        "h2d_bytes_per_step": int(args.molecules * (4 * 8 + 4 + 8)),
        "d2h_bytes_per_step": int(args.size * args.size * 4),
        "delivered_as": "value: Image with the float32 frame in page-locked host memory (float64 array made on demand by "
                        "Image.as_array(), exact); float64_value: Image.as_array() called on every frame "
                        "(widened by the host thread pool, scb_host_widen_*); device_inputs_value: as value, with the "
                        "trajectory resident on the GPU (sample_inputs(..., device=True): no per-frame upload)",
        "frames_timed": n_frames, "warmup_frames": n_warm, "frames_per_block": engine_module.BLOCK_FRAMES,
        "per": "frame (one generate_images iteration; the engine renders blocks of frames_per_block frames)"})
    return out


# --------------------------------------------------------------------------- GPU arm: strong scaling
def run_strong(args):
    """ONE movie of --movie-frames frames (BASELINE.json configs[3]: 10 000) split by frame blocks over
    the ranks.  Timed region per rank: replay of the trajectory + photon budgets of the frames before
    its block (no rendering), then its frames rendered and every block streamed to page-locked host
    memory; max over ranks."""
    import torch
    import torch.distributed as dist
    from scopyon_b200 import _native
    from scopyon_b200.movie import DeviceMovie, frame_block

    world, rank, local, device = setup_dist()
    lib = _native.load()
    T = args.movie_frames
    config = make_config(args.size, args.half_life)
    lower, upper = box(args.size)
    t0 = time.perf_counter()
    movie = DeviceMovie(config, args.molecules, lower, upper, D_COEFF, SEED, device=device, precision="f32")
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    first, last = frame_block(T, rank, world)
    touched = [0.0]

    def sink(first_frame, block):
        touched[0] += float(block[0, 0, 0])

    def one_pass(fmt):
        for _ in range(max(1, min(args.warmup, 2))):          # buffers, pinned ring, clocks
            movie.reset(first_frame=0)
            movie.stream_frames(4 * movie.export_block_frames, fmt=fmt, sink=sink)
        barrier(world)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t_wall = time.perf_counter()
        a.record()
        movie.reset(first_frame=first)
        b.record()
        n = movie.stream_frames(last - first, fmt=fmt, sink=sink)
        wall = time.perf_counter() - t_wall
        c.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(max(a.elapsed_time(c), wall * 1e3), world, device)
        replay = max_over_ranks(a.elapsed_time(b), world, device)
        frames = sum_over_ranks(n, world, device)
        return frames / (ms * 1e-3), ms, replay, frames

    sampler = ClockSampler(local)
    sampler.start()
    rates = {}
    headline = None
    for fmt in ("f32", "u16", "u8"):
        rate, ms, replay, frames = one_pass(fmt)
        rates[fmt] = {"value": rate, "unit": "frames/s", "elapsed_ms": ms, "replay_ms_max_rank": replay,
                      "frames": int(frames)}
        if fmt == args.export:
            headline = (rate, ms, replay)
    clocks = sampler.stop()
    # device-resident rate of the same partition (no export), for comparison
    block = torch.empty((movie.frames_per_launch * 4, args.size, args.size), dtype=torch.float32, device=device)
    movie.reset(first_frame=first)
    ms_res, _, launches, _ = timed_blocks(movie, block, max(1, args.steps), lib, world)
    ms_res = max_over_ranks(ms_res, world, device)
    resident = world * max(1, args.steps) * block.shape[0] / (ms_res * 1e-3)
    host = host_bounds(lib, world, device, args.size * args.size * 4)
    gather_ok = gather_check(args, world, rank, device)
    if rank == 0:
        rate, ms, replay = headline
        n_blocks = (last - first + movie.export_block_frames - 1) // movie.export_block_frames
        line = {
            "metric": METRIC, "value": rate, "unit": "frames/s", "n_gpus": world, "steps": 1, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 box tables and frames, 32-bit fixed-point accumulation per strip (edge arithmetic f64)",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "step": "the whole movie: every rank replays its prefix, renders its frame block and streams every "
                    "finished {}-frame block to page-locked host memory ({}); max over ranks".format(
                        movie.export_block_frames, args.export),
            "replay_ms": replay, "export": rates, "device_resident_frames_per_s": resident,
            "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": int(T * args.size * args.size * {"f32": 4, "u16": 2, "u8": 1}[args.export]),
                    "note": "device-resident trajectory (DeviceMovie API); every frame leaves the GPU inside the timed region"},
            "host": host, "gather_ok": gather_ok,
            "gpu_launches": int(world * n_blocks * (7 if args.export == "f32" else 8)), "setup_s": setup_s,
        }
        if rates["f32"]["value"] and host.get("frames_per_s_at_pcie_f32"):
            line["export"]["f32"]["frac_of_pcie_bound"] = rates["f32"]["value"] / host["frames_per_s_at_pcie_f32"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- GPU arm: C5
def run_c5(args):
    """BASELINE.json configs[4]: 1e6 Gaussian spots per frame on 4096^2.  A step = one frame on resident
    spots: binning + render + detector, on the box-table path (exact) and on the tcgen05 path
    (separable contraction, opt-in, approximate)."""
    import warnings
    import torch
    import scopyon_b200
    from scopyon_b200 import _epifm, _native
    from scopyon_b200.engine import DeviceEngine

    world, rank, local, device = setup_dist()
    if rank != 0:
        return
    size, n = args.size, args.molecules
    config = scopyon_b200.DefaultConfiguration()
    config.update(C5_YAML % (size, size))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(0))
    pl = configs.pixel_length
    rng = numpy.random.RandomState(1)
    data = numpy.zeros((n, 5))
    data[:, 1:3] = rng.uniform(-size * pl / 2, size * pl / 2, (n, 2))
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).to(device)
    weight = torch.full((n,), 30.0, dtype=torch.float64, device=device)
    sw = 2e-9 * (int(round(configs.radial_cutoff / 1e-9)) - 1)
    evals = count_spot_pixel_evals(data, size, pl, sw=sw)
    results, images = {}, {}
    sampler = ClockSampler(local)
    sampler.start()
    for tc in (False, True):
        eng = DeviceEngine(configs, precision="f32", gaussian_tc=tc)
        eng.ensure_tables([0])
        photons = torch.empty((size, size), dtype=torch.float32, device=device)
        adc = torch.empty_like(photons)
        if tc:
            need = eng.lib.scb_gaussian_tc_workspace_bytes(ctypes.byref(eng.geom), n)
            work = torch.empty(need + 256, dtype=torch.uint8, device=device)

            def render():
                eng._call("scb_render_gaussian_tc", ctypes.byref(eng.geom), n, _native.ptr(soa[1]), _native.ptr(soa[2]),
                          _native.ptr(weight), _native.ptr(eng.gaussian_prefix), _native.ptr(photons), _native.F32, 0,
                          _native.ptr(work), work.numel(), _native.ptr(eng.errors), eng._stream())
        else:
            def render():
                eng._render_sat(soa, weight, n, photons, eng._stream())
        for k in range(args.warmup):
            render()
            eng.detect(photons, k, 42, adc=adc)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        render_ms = 0.0
        total_ms = 0.0
        for k in range(args.steps):
            ev[0].record()
            render()
            ev[1].record()
            eng.detect(photons, 100 + k, 42, adc=adc)
            ev[2].record()
            torch.cuda.synchronize()
            render_ms += ev[0].elapsed_time(ev[1])
            total_ms += ev[0].elapsed_time(ev[2])
        images[tc] = photons.double().cpu().numpy()
        results["tcgen05" if tc else "box_table"] = {
            "frames_per_s": args.steps / (total_ms * 1e-3), "render_ms": render_ms / args.steps,
            "frame_ms": total_ms / args.steps, "spot_pixel_evals_per_s": evals / (render_ms / args.steps * 1e-3),
            "errors": int(eng.errors.item())}
    clocks = sampler.stop()
    diff = float(abs(images[True] - images[False]).max() / images[False].max())
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    best = results["box_table"]
    achieved = evals * 4.0 / (best["render_ms"] * 1e-3) / 1e9
    line = {
        "metric": "frames/sec (4096^2, 1e6 Gaussian spots)", "value": best["frames_per_s"], "unit": "frames/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["frame_ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 box tables and frames (box-table path); tf32 hi+lo split operands, fp32 accumulators in TMEM (tcgen05 path)",
        "data": "synthetic",
        "config": {"workload": "C5: separable-Gaussian-PSF stress, {} spots per frame, {}x{} grid, sigma 100 nm, "
                               "resident spots, one frame per step (binning + render + CMOS detector)".format(n, size, size),
                   "molecules": n, "image_size": [size, size],
                   "cache": "one 17 MB box table (L2 resident); 1e6 spot records + unit lists (~0.3 GB) per frame exceed L2"},
        "clocks": clocks, "paths": results, "tcgen05_vs_box_table_max_diff_over_max": diff,
        "roofline": {"kernel": "render pipeline, box-table path", "bound": "issue (table is L2 resident)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "spot_pixel_evals_per_launch": evals},
        "gpu_launches": int(args.steps * 8),
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cpu-check":
        run_cpu_check(args)
    elif args.workload == "C5":
        run_c5(args)
    elif args.scaling == "strong":
        run_strong(args)
    else:
        run_weak(args)


if __name__ == "__main__":
    main()
