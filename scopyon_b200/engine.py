"""Device engine: owns the GPU-resident state of one simulator (PSF summed-area tables,
read-noise alias table, ADC offset map, scratch) and runs one frame through the
C-ABI kernels.  PyTorch is used for device memory, streams and host<->device copies only.
"""
import ctypes
import os

import numpy

from . import _native
from ._epifm import RESOLUTION, catalog_tables, depth_keys_of

try:
    import torch
except ImportError as exc:   # pragma: no cover
    raise ImportError("scopyon_b200 needs PyTorch for device memory management") from exc


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "scopyon_b200 needs a CUDA device (built for NVIDIA B200, sm_100a); there is no CPU fallback")


_libc = ctypes.CDLL(None)
_libc.memcmp.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]


def same_array(a, b):
    """``a`` and ``b`` hold the same values: identity, else one memcmp (no temporaries; an id
    column is compared once per frame on the end-to-end path)."""
    if a is b:
        return True
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if not (a.flags.c_contiguous and b.flags.c_contiguous):
        return bool(numpy.array_equal(a, b))
    return _libc.memcmp(a.ctypes.data, b.ctypes.data, a.nbytes) == 0


# Planes of at least this many pixels take the float32-download + host-widening route (1024 x 1024:
# below that the float64 download is a few tens of microseconds and the thread wake-up dominates).
HOST_WIDEN_MIN_PIXELS = 1 << 20

# Frames enqueued before the oldest is awaited (generate_frames) + 1: plane sets, staging areas.
FRAMES_IN_FLIGHT = 3

# Large float32 frames are handed out as they arrive (image.HostPlane: page-locked float32 payload, float64
# array materialised on demand) while fewer than this many bytes of such payloads are alive; a caller that
# retains more frames than that gets eagerly widened float64 arrays in pageable memory instead.
LAZY_PINNED_BYTES = 1 << 30
#: frames ``generate_frames`` hands to the engine at once when every frame is one snapshot (``begin_block``): binning,
#: rendering and the detector pass then run once per block instead of once per frame.  1 = frame by frame.
BLOCK_FRAMES = int(os.environ.get("SCOPYON_B200_BLOCK_FRAMES", "8"))


def walker_alias(values, weights):
    """Walker/Vose alias table of a categorical distribution -> (n, 4) float32 rows
    (value, alias_value, threshold, 0): the layout of ``scb_alias_entry``."""
    values = numpy.asarray(values, dtype=numpy.float64)
    p = numpy.asarray(weights, dtype=numpy.float64)
    n = len(p)
    scaled = p / p.sum() * n
    threshold = numpy.ones(n)
    alias = numpy.arange(n)
    small = [i for i in range(n) if scaled[i] < 1.0]
    large = [i for i in range(n) if scaled[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        threshold[s] = scaled[s]
        alias[s] = l
        scaled[l] = (scaled[l] + scaled[s]) - 1.0
        (small if scaled[l] < 1.0 else large).append(l)
    table = numpy.zeros((n, 4), dtype=numpy.float32)
    table[:, 0] = values
    table[:, 1] = values[alias]
    table[:, 2] = threshold
    return table


class DeviceBudgetState:
    """Photon budgets per molecule on the device (the reference's ``fluorescence_states``
    dict, ``_epifm.py:1039-1041, 1294-1306``).  NaN marks a molecule not seen yet."""

    def __init__(self, ids, seed, device, initial=None):
        ids = numpy.asarray(ids, dtype=numpy.int64)
        if initial:
            ids = numpy.concatenate([ids, numpy.fromiter(initial.keys(), dtype=numpy.int64, count=len(initial))])
        self.ids = numpy.unique(ids)
        self.seed = int(seed)
        host = numpy.full(len(self.ids), numpy.nan)
        if initial:
            keys = numpy.fromiter(initial.keys(), dtype=numpy.int64, count=len(initial))
            host[numpy.searchsorted(self.ids, keys)] = numpy.fromiter(
                initial.values(), dtype=numpy.float64, count=len(initial))
        self.budget = torch.from_numpy(host).to(device)

    def slots_of(self, ids):
        ids = numpy.asarray(ids, dtype=numpy.int64)
        slots = numpy.searchsorted(self.ids, ids)
        slots = numpy.minimum(slots, len(self.ids) - 1)
        if len(ids) and not numpy.array_equal(self.ids[slots], ids):
            raise ValueError("a molecule id absent from the input data was given")
        return slots.astype(numpy.int32)

    def as_dict(self, only_seen=True):
        host = self.budget.cpu().numpy()
        keep = ~numpy.isnan(host) if only_seen else numpy.ones(len(host), dtype=bool)
        return {int(i): float(b) for i, b in zip(self.ids[keep], host[keep])}


class SatStore:
    """Device-resident PSF summed-area tables of one PSF model, grown on demand.

    ``sat[slot]`` is the int64 table of one depth key (``scb_psf_sat_build``),
    ``slot_of_key`` maps depth keys to slots (-1: not built).  All depth keys of the
    default configuration take 1003 x 32 MB = 32 GB of the B200's 180 GB."""

    _shared = {}

    @classmethod
    def shared(cls, engine):
        cfg = engine.configs
        key = (str(engine.device), engine.psf_type, float(cfg.psf_wavelength), cfg.psf_radial_width,
               engine.geom.n_radial, engine.geom.n_depth_keys, float(cfg.depth_cutoff), engine.geom.sat_modulus,
               cls.wants_box(engine), engine.precision)
        store = cls._shared.get(key)
        if store is None:
            store = cls._shared[key] = cls(engine)
        return store

    @classmethod
    def clear_shared(cls):
        cls._shared.clear()

    @staticmethod
    def wants_box(engine):
        """Box tables (the per-pixel integrals the TMA render path reads) only pay off when the
        pixel pitch is a whole number of table samples: otherwise no footprint has evenly
        spaced edges and every spot is summed from the SAT corners."""
        ratio = float(engine.geom.pixel_length) / float(engine.geom.resolution)
        return abs(ratio - round(ratio)) < 1e-6 * ratio

    def __init__(self, engine):
        cfg = engine.configs
        self.lib = engine.lib
        self.device = engine.device
        self.psf_type = engine.psf_type
        self.wavelength = float(cfg.psf_wavelength)
        self.radial_width = float(cfg.psf_radial_width or 0.0)
        self.depth_cutoff = float(cfg.depth_cutoff)
        self.n_radial = int(engine.geom.n_radial)
        self.n_depth_keys = int(engine.geom.n_depth_keys)
        self.modulus = int(engine.geom.sat_modulus)
        self.rows = 2 * (self.n_radial - 1) + 2
        self.slots = int(self.lib.scb_psf_sat_slots(self.n_radial, self.modulus))
        self.table_entries = int(self.lib.scb_psf_sat_table_entries(self.n_radial, self.modulus))
        self.with_box = self.wants_box(engine)
        # box values in the precision of the frames: fp32 tables halve the memory and the DRAM
        # traffic of the render (6e-8 relative per pixel, the rounding fp32 frames have anyway)
        self.box_dtype = torch.float32 if engine.precision == "f32" else torch.float64
        self.box_type = _native.F32 if engine.precision == "f32" else _native.F64
        self.sat = None
        self.box = None
        self.box_peak = 0.0      # largest fraction of a spot's photons one pixel receives (over the tables built so far)
        self.inv_scale = None
        self.n_tables = 0
        self.last_radial = None
        self.slot_host = numpy.full(self.n_depth_keys + 1, -1, dtype=numpy.int32)
        self.slot_of_key = torch.from_numpy(self.slot_host.copy()).to(self.device)

    def table_depth(self, key):
        """Depth a table is evaluated at (``_epifm.py:80-84``)."""
        return float(key) * RESOLUTION if key < self.n_depth_keys else self.depth_cutoff

    def index_of(self, a, b):
        """Position of ``S[a][b]`` inside a stored table (the block layout of
        ``csrc/scb_common.cuh``): block ``(a % M, b % M)``, slot ``(a // M, b // M)``."""
        m, s = self.modulus, self.slots
        return (((a % m) * m + (b % m)) * s + a // m) * s + b // m

    def plain(self, slot):
        """Table ``slot`` on the host in plain ``S[a][b]`` order (undoes the block layout)."""
        stored = self.sat[slot].cpu().numpy()
        a = numpy.arange(self.rows)
        return stored[self.index_of(a[:, None], a[None, :])]

    def all_resident(self):
        return self.n_tables > 0 and bool((self.slot_host >= 0).all())

    def ensure(self, keys, stream):
        """Build the tables of depth keys not resident yet."""
        keys = numpy.unique(numpy.asarray(keys, dtype=numpy.int64))
        if self.psf_type == _native.PSF_GAUSSIAN:
            # depth independent (_epifm.py:133-134): one table serves every key
            if self.n_tables == 0:
                self.build([0], stream)
                self.slot_host[:] = 0
                self.slot_of_key.copy_(torch.from_numpy(self.slot_host))
            return
        missing = [int(k) for k in keys if self.slot_host[k] < 0]
        if not missing:
            return
        first = self.build(missing, stream)
        for i, k in enumerate(missing):
            self.slot_host[k] = first + i
        self.slot_of_key.copy_(torch.from_numpy(self.slot_host))

    def build(self, keys, stream, radial=None):
        """Append tables for ``keys``; returns the first new slot.  ``radial`` (n, n_radial)
        overrides the device-computed radial profiles (parity tests feed the oracle's)."""
        n_new = len(keys)
        need = self.n_tables + n_new
        capacity = 0 if self.sat is None else self.sat.shape[0]
        if need > capacity:
            new_cap = max(min(max(need, 2 * capacity, 1), self.n_depth_keys + 1), need)
            sat = torch.empty((new_cap, self.table_entries), dtype=torch.int64, device=self.device)
            box = torch.empty((new_cap, self.table_entries), dtype=self.box_dtype, device=self.device) \
                if self.with_box else None
            inv = torch.zeros(new_cap, dtype=torch.float64, device=self.device)
            if self.n_tables:
                sat[:self.n_tables].copy_(self.sat[:self.n_tables])
                inv[:self.n_tables].copy_(self.inv_scale[:self.n_tables])
                if box is not None:
                    box[:self.n_tables].copy_(self.box[:self.n_tables])
            self.sat, self.box, self.inv_scale = sat, box, inv
        if radial is None:
            depths = torch.tensor([self.table_depth(k) for k in keys], dtype=torch.float64).to(self.device)
            radial = torch.empty((n_new, self.n_radial), dtype=torch.float64, device=self.device)
            _native.check(self.lib.scb_psf_radial_build(
                self.psf_type, self.wavelength, self.radial_width, self.n_radial, n_new,
                _native.ptr(depths), _native.ptr(radial), stream), "scb_psf_radial_build")
        work_bytes = self.lib.scb_psf_sat_workspace_bytes(self.n_radial, n_new)
        work = torch.empty(work_bytes, dtype=torch.uint8, device=self.device)
        first = self.n_tables
        _native.check(self.lib.scb_psf_sat_build(
            _native.ptr(radial), self.n_radial, n_new, self.modulus, ctypes.c_void_p(self.sat[first].data_ptr()),
            None if self.box is None else ctypes.c_void_p(self.box[first].data_ptr()), self.box_type,
            ctypes.c_void_p(self.inv_scale[first:].data_ptr()), _native.ptr(work), work_bytes, stream),
            "scb_psf_sat_build")
        if self.box is not None:
            # sizes the 32-bit accumulators of the fp32 render (scb_geometry.box_peak)
            peaks = self.box[first:need].amax(dim=1).to(torch.float64) * self.inv_scale[first:need] * (RESOLUTION ** 2)
            self.box_peak = max(self.box_peak, float(peaks.max().item()))
        self.n_tables = need
        self.last_radial = radial
        return first


class _Trace:
    """Optional host-side stage timer (``engine.trace = {}`` enables it; bench.py reports it)."""

    def __init__(self, engine, name, sync=False):
        self.engine, self.name, self.sync = engine, name, sync

    def __enter__(self):
        import time
        self.t0 = time.perf_counter() if self.engine.trace is not None else None

    def __exit__(self, *exc):
        if self.t0 is not None:
            import time
            if self.sync:
                torch.cuda.current_stream(self.engine.device).synchronize()
            acc = self.engine.trace
            acc[self.name] = acc.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3


class DeviceEngine:

    trace = None
    _defer_true_data = False

    def __init__(self, configs, device=None, precision=None, gaussian_tc=None):
        require_cuda()
        self.lib = _native.load()
        self.configs = configs
        self.device = torch.device(device if device is not None else "cuda:{}".format(torch.cuda.current_device()))
        self.precision = precision or "f32"
        if self.precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        self.elem_type = _native.F32 if self.precision == "f32" else _native.F64
        self.dtype = torch.float32 if self.precision == "f32" else torch.float64
        self._geom = configs.geometry()
        self.phys = configs.photophysics()
        self.det = configs.detector_struct()
        self.psf_type = _native.PSF_GAUSSIAN if configs.fluorophore_type == 'Gaussian' else _native.PSF_BORN_WOLF
        if self.psf_type == _native.PSF_GAUSSIAN and configs.psf_radial_width is None:
            raise ValueError('fluorophore.radial_width must be given for Gaussian type fluorophore.')
        self.n_w, self.n_h = int(self.geom.n_w), int(self.geom.n_h)
        # opt-in tensor-core renderer for the separable Gaussian PSF (1e-5 of the image maximum
        # instead of exact; see csrc/gaussian_tc.cu)
        if gaussian_tc is None:
            import os
            gaussian_tc = os.environ.get("SCOPYON_B200_GAUSSIAN_TC", "0") == "1"
        self.gaussian_tc = bool(gaussian_tc) and self.psf_type == _native.PSF_GAUSSIAN
        self.gaussian_prefix = None

        # PSF summed-area tables: shared by every engine of this process with the same PSF
        # (the reference rebuilds its table cache on each form_image call, base.py:56-59)
        self.tables = SatStore.shared(self)
        self.errors = torch.zeros(1, dtype=torch.int32, device=self.device)
        if self.gaussian_tc:
            # 1-D factor of the separable form = the marginal of the reference's own Cartesian table
            # (row sums), whose running sum is the last column of the summed-area table:
            #   T[a][b] ~= m(a) m(b) / mass,  prefix[a] = sum_{a' < a} m(a') / sqrt(mass).
            # Against the reference table this is accurate to 1.1e-6 of the peak at sigma = 100 nm
            # (the analytic Gaussian: 7e-6) and has no bias in dense fields.
            self.ensure_tables([0])
            side = self.tables.rows - 1
            last_column = torch.from_numpy(self.tables.index_of(numpy.arange(side + 1), side)).to(self.device)
            cumulative = self.tables.sat[0][last_column].to(torch.float64) * self.tables.inv_scale[0] * (RESOLUTION ** 2)
            self.gaussian_prefix = (cumulative / torch.sqrt(cumulative[-1])).contiguous()

        # detector-side constants
        self.alias = None
        if configs.detector_type == "CMOS":
            rn = catalog_tables()["cmos_readout"]
            self.alias = torch.from_numpy(walker_alias(rn["electrons"], rn["weight"])).to(self.device)
        self.offset = None
        fpn = configs.ADConverter_fpn_type
        if fpn != 'none':
            n = self.n_w * self.n_h if fpn == 'pixel' else self.n_h
            self.offset = torch.empty(n, dtype=self.dtype, device=self.device)
            self._call("scb_adc_offsets", configs.fpn_seed, n, float(configs.ADConverter_offset0),
                       float(configs.ADConverter_fpn_count), _native.ptr(self.offset), self.elem_type,
                       self._stream())
        self.max_pinned_planes = 16
        self._soa_cols = torch.tensor([0, 1, 2, 4], device=self.device)
        self._workspace = None
        self._det_work = torch.empty(self.lib.scb_detector_workspace_bytes(self.n_w, self.n_h), dtype=torch.uint8,
                                     device=self.device)
        self._expected = torch.zeros((self.n_w, self.n_h), dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ plumbing
    @property
    def geom(self):
        """The ``scb_geometry`` of this simulator; ``box_peak`` follows the shared table store
        (tables may have been added by another simulator since the last call)."""
        tables = getattr(self, "tables", None)
        self._geom.box_peak = tables.box_peak if tables is not None else 0.0
        return self._geom

    def _stream(self):
        cached = getattr(self, "_frame_stream", None)      # set for the duration of begin_frame
        if cached is not None:
            return cached
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _call(self, name, *args):
        # the library launches on the current CUDA device: make it this engine's for the call
        with torch.cuda.device(self.device):
            _native.check(getattr(self.lib, name)(*args), name)

    def _to_device(self, array, dtype=None):
        host = torch.from_numpy(numpy.ascontiguousarray(array))
        if dtype is not None:
            host = host.to(dtype)
        return host.pin_memory().to(self.device, non_blocking=True)

    def _h2d_stage(self, n):
        """Pinned (capacity, 5) float64 staging area for particle uploads.  ``FRAMES_IN_FLIGHT``
        areas rotate: with the lookahead of ``generate_frames`` the uploads of the previous frames
        may still be queued while the host already fills the area for the next one."""
        stages = getattr(self, "_stage_h2d", None)
        if stages is None:
            stages = self._stage_h2d = [None] * FRAMES_IN_FLIGHT
        stage = stages[self._stage_turn]
        if stage is None or stage.shape[0] < n:
            capacity = max(n, 2 * (0 if stage is None else stage.shape[0]), 1024)
            stage = stages[self._stage_turn] = torch.empty((capacity, 5), dtype=torch.float64).pin_memory()
        return stage

    def _host_plane(self, shape=None):
        """A float64 (Nw, Nh) host array for one returned plane.  While few planes are alive
        (streaming use) it is pinned memory the device writes directly, so no host copy
        is needed; when the caller retains many frames it falls back to pageable memory."""
        import weakref
        shape = (self.n_w, self.n_h) if shape is None else tuple(shape)
        live = getattr(self, "_live_planes", None)
        if live is None:
            live = self._live_planes = weakref.WeakSet()
        if len(live) < self.max_pinned_planes:
            t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
            live.add(t)
            return t, True
        return torch.empty(shape, dtype=torch.float64), False

    def _render_workspace(self, n_spots):
        need = self.lib.scb_render_workspace_bytes(ctypes.byref(self.geom), n_spots)
        if need == 0:
            raise _native.NativeError("scb_render_workspace_bytes: " + self.lib.scb_last_error().decode())
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return self._workspace

    # ------------------------------------------------------------------ PSF tables
    sat = property(lambda self: self.tables.sat)
    box = property(lambda self: self.tables.box)
    box_type = property(lambda self: self.tables.box_type)
    inv_scale = property(lambda self: self.tables.inv_scale)
    slot_of_key = property(lambda self: self.tables.slot_of_key)
    slot_host = property(lambda self: self.tables.slot_host)
    n_tables = property(lambda self: self.tables.n_tables)
    last_radial = property(lambda self: self.tables.last_radial)

    def table_depth(self, key):
        return self.tables.table_depth(key)

    def ensure_tables(self, keys):
        self.tables.ensure(keys, self._stream())

    def ensure_all_tables(self):
        self.tables.ensure(numpy.arange(self.geom.n_depth_keys + 1), self._stream())

    def _build_tables(self, keys, radial=None):
        return self.tables.build(keys, self._stream(), radial=radial)

    # ------------------------------------------------------------------ states
    def new_budget_state(self, input_data, seed, initial=None):
        def ids_of(p):
            ids = getattr(p, "ids", None)
            return ids if ids is not None and len(ids) == len(p) else numpy.asarray(numpy.asarray(p)[:, 3], dtype=numpy.int64)
        columns, seen = [], set()
        for _, p in input_data:
            ids = ids_of(p)
            if id(ids) not in seen:         # snapshots of one trajectory share their id column
                seen.add(id(ids))
                columns.append(ids)
        ids = numpy.concatenate(columns) if columns else numpy.zeros(0, dtype=numpy.int64)
        return DeviceBudgetState(ids, seed, self.device, initial=initial)

    # ------------------------------------------------------------------ one frame
    def render_expected(self, snapshots, states=None, want_true_data=False, out=None, exposure_time=None):
        """Expected photon image (device tensor) of the given ``[(unit_time, particles(N,5))]``
        plus the post-processed ``true_data`` dict (or None)."""
        cfg = self.configs
        out = self._expected if out is None else out
        sizes = [len(p) for _, p in snapshots]
        total = int(sum(sizes))
        stream = self._stream()
        focal = cfg.detector_focal_point
        true_dev, true_ids = None, None
        if total == 0:
            out.zero_()
            return out, ({} if want_true_data else None)

        with _Trace(self, "host_prepare"):
            return self._render_expected(snapshots, states, want_true_data, out, exposure_time, sizes, total)

    def _render_expected(self, snapshots, states, want_true_data, out, exposure_time, sizes, total):
        cfg = self.configs
        stream = self._stream()
        focal = cfg.detector_focal_point
        true_dev, true_ids = None, None
        on_device = [hasattr(p, "tensor") for _, p in snapshots]     # base.DeviceRows: the trajectory never left the GPU
        lone = None
        if len(snapshots) == 1 and on_device[0] and snapshots[0][1].tensor.device == self.device:
            lone = snapshots[0][1].tensor                              # rendered where it lies
        rows_dev = lone if lone is not None else torch.empty((total, 5), dtype=torch.float64, device=self.device)
        weight = torch.empty(total, dtype=torch.float64, device=self.device)
        self._stage_turn = (getattr(self, "_stage_turn", 0) + 1) % FRAMES_IN_FLIGHT   # this frame's staging area (see _h2d_stage)
        all_ids = self._ids_of(snapshots)
        if states is not None:
            table_ids = states.ids
        elif want_true_data:
            table_ids = numpy.unique(all_ids)
        else:
            table_ids = None
        if want_true_data:
            true_ids = table_ids
            true_dev = torch.zeros((len(table_ids), 8), dtype=torch.float64, device=self.device)

        keys = []
        offset = 0
        all_resident = self.tables.all_resident()
        for (unit_time, particles), n, resident in zip(snapshots, sizes, on_device):
            if n == 0:
                continue
            ids = all_ids if n == total else all_ids[offset: offset + n]    # the cached object itself when it can be
            order, rounds, slots_dev = None, [(0, n)], None
            if table_ids is not None:
                order, rounds, slots_dev, _ = self._molecule_slots(table_ids, ids)
            if resident:
                source = particles.tensor
                if order is not None:
                    source = source[torch.from_numpy(order).to(self.device)]
                    rows_dev = rows_dev if lone is None else torch.empty((total, 5), dtype=torch.float64, device=self.device)
                    lone = None
                if lone is None:
                    rows_dev[offset: offset + n].copy_(source, non_blocking=True)
                if not all_resident:
                    depth = (source[:, 0] - float(focal[0])).abs()
                    key = torch.clamp((depth / RESOLUTION).to(torch.int64), max=self.geom.n_depth_keys - 1)
                    key = torch.where(depth < cfg.depth_cutoff + RESOLUTION, key,
                                      torch.full_like(key, self.geom.n_depth_keys))
                    keys.append(torch.unique(key).cpu().numpy())
            else:
                # rows go up as they are: straight from the caller's array when that is page-locked
                # (base.__format_data allocates it so), else through a pinned staging copy
                host = None
                if order is None and isinstance(particles, numpy.ndarray) and particles.dtype == numpy.float64 \
                        and particles.flags.c_contiguous:
                    host = torch.from_numpy(particles.view(numpy.ndarray))
                    if not host.is_pinned():
                        host = None
                if host is None:
                    particles = numpy.asarray(particles, dtype=numpy.float64)
                    if order is not None:
                        particles = particles[order]
                    host = self._h2d_stage(total)[offset: offset + n]
                    numpy.copyto(host.numpy(), particles)
                rows_dev[offset: offset + n].copy_(host, non_blocking=True)
                if not all_resident:
                    keys.append(depth_keys_of(numpy.asarray(particles)[:, 0] - focal[0], cfg.depth_cutoff,
                                              self.geom.n_depth_keys))
            for lo, hi in rounds:
                self._call(
                    "scb_emit_bleach_rows", states.seed if states is not None else 0, hi - lo,
                    _native.ptr(rows_dev[offset + lo: offset + hi]),
                    None if slots_dev is None else _native.ptr(slots_dev[lo:hi]),
                    float(unit_time), float(focal[0]), ctypes.byref(self.phys),
                    None if states is None else _native.ptr(states.budget),
                    _native.ptr(weight[offset + lo: offset + hi]),
                    None if true_dev is None else _native.ptr(true_dev), stream)
            offset += n

        if self.gaussian_tc:
            need = self.lib.scb_gaussian_tc_workspace_bytes(ctypes.byref(self.geom), total)
            if self._workspace is None or self._workspace.numel() < need:
                self._workspace = torch.empty(int(need * 1.25) + 256, dtype=torch.uint8, device=self.device)
            xy = rows_dev[:, 1:3].t().contiguous()
            self._call(
                "scb_render_gaussian_tc", ctypes.byref(self.geom), total, _native.ptr(xy[0]), _native.ptr(xy[1]),
                _native.ptr(weight), _native.ptr(self.gaussian_prefix), _native.ptr(out),
                _native.F32 if out.dtype == torch.float32 else _native.F64, 0, _native.ptr(self._workspace),
                self._workspace.numel(), _native.ptr(self.errors), stream)
            keys = None
        elif keys:
            needed = numpy.unique(numpy.concatenate(keys))
            # a 3-D scene that touches many depth keys will touch all of them soon: build the lot
            if self.psf_type != _native.PSF_GAUSSIAN and (self.slot_host[needed] < 0).sum() > 128:
                self.ensure_all_tables()
            else:
                self.ensure_tables(needed)
        if not self.gaussian_tc:
            work = self._render_workspace(total)
            self._call(
                "scb_render_expected_rows", ctypes.byref(self.geom), total, _native.ptr(rows_dev), _native.ptr(weight),
                _native.ptr(self.sat), _native.ptr(self.box), self.box_type, _native.ptr(self.inv_scale), _native.ptr(self.slot_of_key),
                _native.ptr(out), _native.F32 if out.dtype == torch.float32 else _native.F64, 0,
                _native.ptr(work), work.numel(), _native.ptr(self.errors), stream)

        if self._defer_true_data:
            return out, (true_dev, true_ids) if want_true_data else None
        true_data = None
        if want_true_data:
            true_data = self._finish_true_data(true_dev.cpu().numpy(), true_ids, exposure_time)
        return out, true_data

    def _ids_of(self, snapshots):
        """int64 molecule ids of all snapshot rows.  Rows formatted by the facade carry them as a
        contiguous column (``ParticleRows.ids``); consecutive frames usually show the same ids, in
        which case the previous frame's array object is handed out again (the slot cache keys on it)."""
        cols = []
        for _, p in snapshots:
            ids = getattr(p, "ids", None)
            if ids is None or len(ids) != len(p):
                ids = numpy.ascontiguousarray(numpy.asarray(p)[:, 3]).astype(numpy.int64)
            cols.append(ids)
        cache = getattr(self, "_ids_cache", None)
        if cache is not None and len(cache[0]) == len(cols) and all(same_array(a, b) for a, b in zip(cache[0], cols)):
            return cache[1]
        ids = numpy.concatenate(cols) if cols else numpy.zeros(0, numpy.int64)
        self._ids_cache = (cols, ids)
        return ids

    def _render_sat(self, soa, weight, total, out, stream):
        work = self._render_workspace(total)
        self._call(
            "scb_render_expected", ctypes.byref(self.geom), total,
            _native.ptr(soa[0]), _native.ptr(soa[1]), _native.ptr(soa[2]), _native.ptr(weight),
            _native.ptr(self.sat), _native.ptr(self.box), self.box_type, _native.ptr(self.inv_scale), _native.ptr(self.slot_of_key),
            _native.ptr(out), _native.F32 if out.dtype == torch.float32 else _native.F64, 0,
            _native.ptr(work), work.numel(), _native.ptr(self.errors), stream)

    def _molecule_slots(self, table_ids, ids):
        """Device copies of (budget/true_data slot, molecule id) per particle row, cached while
        consecutive snapshots carry the same id column (the usual case).  When an id repeats
        inside one snapshot the reference updates its budget row by row, so the rows are
        reordered into rounds (occurrence k launched after occurrence k-1)."""
        cache = getattr(self, "_slot_cache", None)
        if cache is not None and cache[0] is table_ids and same_array(cache[1], ids):
            return cache[2]
        n = len(ids)
        order, rounds = None, [(0, n)]
        uniq, counts = numpy.unique(ids, return_counts=True)
        if len(uniq) < n:
            by_id = numpy.argsort(ids, kind='stable')
            rank = numpy.empty(n, dtype=numpy.int64)
            starts = numpy.concatenate([[0], numpy.cumsum(counts)[:-1]])
            rank[by_id] = numpy.arange(n) - numpy.repeat(starts, counts)
            order = numpy.argsort(rank, kind='stable')
            bounds = numpy.concatenate([[0], numpy.cumsum(numpy.bincount(rank))])
            rounds = [(int(bounds[i]), int(bounds[i + 1])) for i in range(len(bounds) - 1)]
        ordered = ids if order is None else ids[order]
        slots = numpy.searchsorted(table_ids, ordered)
        # an id outside the table would alias another molecule's budget (or write past the end)
        if n and (len(table_ids) == 0 or not numpy.array_equal(
                table_ids[numpy.minimum(slots, len(table_ids) - 1)], ordered)):
            raise ValueError("a molecule id absent from the fluorescence states / input data was given")
        slots = slots.astype(numpy.int32)
        result = (order, rounds, self._to_device(slots), self._to_device(ordered))
        self._slot_cache = (table_ids, ids, result)
        return result

    def _finish_true_data(self, acc, ids, exposure_time):
        """Time averages and pixel coordinates of the per-molecule vector (``_epifm.py:1207-1216``)."""
        cfg = self.configs
        p_0 = cfg.detector_focal_point
        pl = cfg.pixel_length
        seen = acc[:, 0] > 0
        acc, ids = acc[seen], ids[seen]
        acc[:, 1] /= exposure_time
        acc[:, 2:6] /= acc[:, 0:1]
        acc[:, 2] = (acc[:, 2] - p_0[1]) / pl + (self.n_w - 1) * 0.5
        acc[:, 3] = (acc[:, 3] - p_0[2]) / pl + (self.n_h - 1) * 0.5
        return {int(i): row.copy() for i, row in zip(ids, acc)}

    def detect(self, photons, frame_index, noise_seed, adc=None, expectation=None,
               in_signal=None, in_noise=None, out_signal=None, out_noise=None):
        """Detector + ADC pass on a device image (``_epifm.py:1430-1470``)."""
        if adc is None:
            adc = torch.empty_like(photons)
        elem = _native.F32 if photons.dtype == torch.float32 else _native.F64
        offset = self.offset
        if offset is not None and offset.dtype != photons.dtype:
            offset = offset.to(photons.dtype)
        self._call(
            "scb_detector_adc", int(noise_seed), int(frame_index), ctypes.byref(self.det),
            self.n_w, self.n_h, elem, _native.ptr(photons), _native.ptr(offset),
            _native.ptr(self.alias), 0 if self.alias is None else int(self.alias.shape[0]),
            _native.ptr(adc), _native.ptr(expectation), _native.ptr(in_signal), _native.ptr(in_noise),
            _native.ptr(out_signal), _native.ptr(out_noise), _native.ptr(self._det_work), self._det_work.numel(),
            self._stream())
        return adc

    def begin_frame(self, snapshots, frame_index, noise_seed, states, exposure_time, want_true_data,
                    want_expectation=True, snapshot_states=False, lazy=True):
        """Enqueue one frame (upload, kernels, download into pinned host planes) and return a
        handle without waiting for the device; ``finish_frame`` waits and hands out the arrays.
        Small planes are widened to float64 on the device and written straight into pinned host
        arrays (one DMA per plane); large float32 planes are downloaded as they are -- into the
        page-locked payload of an ``image.HostPlane`` when ``lazy`` (the float64 array is then made
        on demand), else into staging memory that host threads widen while the next frames run."""
        with torch.cuda.device(self.device):        # kernels launch on the current device: make it ours
            main = torch.cuda.current_stream(self.device)
            self._frame_stream = ctypes.c_void_p(main.cuda_stream)      # one stream lookup per frame
            try:
                return self._begin_frame(main, snapshots, frame_index, noise_seed, states, exposure_time,
                                         want_true_data, want_expectation, snapshot_states, lazy)
            finally:
                self._frame_stream = None

    def _widen_setup(self):
        """Worker threads of the widening pool: the ranks of one box (one process per GPU) share its host
        cores, so each takes its share of them and pins its workers there."""
        if getattr(self, "_widen_ready", False):
            return
        self._widen_ready = True
        ranks_here = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        if ranks_here <= 1:
            return
        local = int(os.environ.get("LOCAL_RANK", "0") or 0)
        try:
            cpus = sorted(os.sched_getaffinity(0))
        except AttributeError:      # pragma: no cover
            cpus = list(range(os.cpu_count() or 1))
        share = max(1, len(cpus) // ranks_here)
        mine = cpus[local * share: (local + 1) * share] or cpus
        # leave one core of the share to the interpreter thread when there is a choice
        workers = mine[1:] if len(mine) > 2 else mine
        self.lib.scb_host_widen_threads(max(1, min(6, len(workers))))
        self.lib.scb_host_widen_affinity((ctypes.c_int * len(workers))(*workers), len(workers))

    def widen_host(self, f32):
        """float32 host array -> the float64 array of the API, on the widening pool (exact)."""
        self._widen_setup()
        dst, _ = self._host_plane(shape=f32.shape)
        src = numpy.ascontiguousarray(f32)
        ticket = self.lib.scb_host_widen_start(src.ctypes.data, dst.data_ptr(), src.size, None, 0)
        if ticket <= 0:
            raise _native.NativeError("scb_host_widen_start: " + self.lib.scb_last_error().decode())
        _native.check(self.lib.scb_host_widen_wait(ticket), "scb_host_widen_wait")
        return dst.numpy()

    def _lazy_plane(self):
        """A page-locked float32 (Nw, Nh) payload for one frame, or None when the caller already holds
        ``LAZY_PINNED_BYTES`` of them.  Payloads are pooled: ``_recycle`` hands one back once neither the
        frame nor any numpy view of it is alive, so a streaming consumer never pays for page-locking again."""
        pool = self.__dict__.setdefault("_lazy_pool", dict(free=[], allocated=0))
        if pool["free"]:
            return pool["free"].pop()
        if (pool["allocated"] + 1) * self.n_w * self.n_h * 4 > LAZY_PINNED_BYTES:
            return None
        pool["allocated"] += 1
        return torch.empty((self.n_w, self.n_h), dtype=torch.float32, pin_memory=True)

    def _lazy_view(self, tensor):
        """numpy view of a pooled payload; the tensor returns to the pool when the view (and with it every
        view derived from it: they keep it alive as their ``base``) has been collected."""
        import weakref
        array = tensor.numpy()
        weakref.finalize(array, self._lazy_pool["free"].append, tensor)
        return array

    def _float64_wanted(self, f32):
        """``widen`` callback of the HostPlanes this engine hands out: the caller asks for float64 arrays, so
        the following frames are widened while they are in flight (see ``begin_frame``) instead of on demand."""
        self._eager_float64 = True
        return self.widen_host(f32)

    def _begin_frame(self, main, snapshots, frame_index, noise_seed, states, exposure_time, want_true_data,
                     want_expectation, snapshot_states, lazy=True):
        self._defer_true_data = True
        try:
            photons, true_pending = self.render_expected(
                snapshots, states=states, want_true_data=want_true_data, exposure_time=exposure_time)
        finally:
            self._defer_true_data = False
        if getattr(self, "_planes32", None) is None:
            # FRAMES_IN_FLIGHT plane sets rotate: the download of frame f (copy stream) overlaps the
            # kernels of the frames behind it; generate_frames awaits frame f before it starts
            # frame f + FRAMES_IN_FLIGHT
            self._planes32 = torch.empty((FRAMES_IN_FLIGHT, 2, self.n_w, self.n_h), dtype=self.dtype, device=self.device)
            self._planes_free = [None] * FRAMES_IN_FLIGHT       # download of the set's previous frame
            self._stage32 = self._planes64 = None
            self._stage_tickets = [[] for _ in range(FRAMES_IN_FLIGHT)]
            self._large32 = self.dtype == torch.float32 and self.n_w * self.n_h >= HOST_WIDEN_MIN_PIXELS
            if self.dtype == torch.float32 and not self._large32:
                # small planes (the wake-up of the host threads would cost more than the bytes saved)
                # are widened on the device and downloaded as float64
                self._planes64 = torch.empty((FRAMES_IN_FLIGHT, 2, self.n_w, self.n_h), dtype=torch.float64,
                                             device=self.device)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._plane_turn = 0
        turn = self._plane_turn = (self._plane_turn + 1) % FRAMES_IN_FLIGHT
        if self._planes_free[turn] is not None:
            main.wait_event(self._planes_free[turn])            # the set's previous frame has left the device
        for ticket in self._stage_tickets[turn]:                # ... and its staging area has been read
            _native.check(self.lib.scb_host_widen_wait(ticket), "scb_host_widen_wait")
        self._stage_tickets[turn] = []
        p32 = self._planes32[turn]
        self.detect(photons, frame_index, noise_seed, adc=p32[0], expectation=p32[1] if want_expectation else None)
        n_planes = 2 if want_expectation else 1
        payloads = None
        with _Trace(self, "host_alloc_planes"):
            if self._large32 and lazy and not getattr(self, "_eager_float64", False):
                payloads = [self._lazy_plane() for _ in range(n_planes)]
                if any(p is None for p in payloads):
                    self._lazy_pool["free"].extend(p for p in payloads if p is not None)
                    payloads = None
            if payloads is None:
                hosts = [self._host_plane()[0] for _ in range(n_planes)]
        staged = self._large32 and payloads is None
        if staged and self._stage32 is None:
            # eager route for large float32 planes: pinned staging memory, widened by host threads
            self._stage32 = torch.empty((FRAMES_IN_FLIGHT, 2, self.n_w, self.n_h), dtype=torch.float32, pin_memory=True)
            self._widen_setup()
        tickets = []
        with _Trace(self, "enqueue_d2h"):
            sources = p32
            if self._planes64 is not None:
                sources = self._planes64[turn]
                for k in range(n_planes):
                    sources[k].copy_(p32[k])
            ready = torch.cuda.Event()
            ready.record(main)
            self._copy_stream.wait_event(ready)
            with torch.cuda.stream(self._copy_stream):
                if payloads is not None:
                    for k, host in enumerate(payloads):
                        host.copy_(p32[k], non_blocking=True)       # float32 planes: the frame's own payload
                elif not staged:
                    for k, host in enumerate(hosts):
                        host.copy_(sources[k], non_blocking=True)  # float64 planes: straight into the caller's array
                else:
                    for k in range(n_planes):
                        self._stage32[turn][k].copy_(p32[k], non_blocking=True)
                planes_done = torch.cuda.Event()
                planes_done.record(self._copy_stream)
            self._planes_free[turn] = planes_done
            if staged:
                device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
                for k, host in enumerate(hosts):
                    ticket = self.lib.scb_host_widen_start(
                        self._stage32[turn][k].data_ptr(), host.data_ptr(), host.numel(),
                        ctypes.c_void_p(planes_done.cuda_event), device_index)
                    if ticket <= 0:
                        raise _native.NativeError("scb_host_widen_start: " + self.lib.scb_last_error().decode())
                    tickets.append(ticket)
                self._stage_tickets[turn] = tickets
        true_host = None
        if isinstance(true_pending, tuple):
            true_host = (torch.empty(true_pending[0].shape, dtype=torch.float64, pin_memory=True), true_pending[1])
            true_host[0].copy_(true_pending[0], non_blocking=True)
        elif true_pending is not None:
            true_host = true_pending                       # the empty-frame dict
        budget_host = None
        if snapshot_states and states is not None:
            budget_host = torch.empty(states.budget.shape, dtype=torch.float64, pin_memory=True)
            budget_host.copy_(states.budget, non_blocking=True)
        # this frame's count of spots without a PSF table; the counter restarts for the next frame, so one
        # failed frame is reported once and not again by the frames already in flight behind it
        errors_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        errors_host.copy_(self.errors, non_blocking=True)
        self.errors.zero_()
        done = torch.cuda.Event()
        done.record(main)
        return dict(hosts=payloads if payloads is not None else hosts, lazy=payloads is not None,
                    true=true_host, budget=budget_host, states=states, done=done, errors=errors_host,
                    planes_done=planes_done, tickets=tickets,
                    exposure_time=exposure_time, want_expectation=want_expectation)

    # ------------------------------------------------------------------ blocks of frames (generate_frames)
    def block_route(self, want_full_output):
        """Whether ``generate_frames`` may hand this engine several frames at once (``begin_block``): bare ADC
        planes of large float32 frames -- the case where a movie is made of many frames of one snapshot each."""
        if not (BLOCK_FRAMES > 1 and not want_full_output and not self.gaussian_tc and self.dtype == torch.float32
                and (self.n_w * self.n_h) % 4 == 0
                and (self.configs.ADConverter_fpn_type != 'column' or self.n_h % 4 == 0)):
            return False
        # large frames: float32 payloads (not once the caller asks for float64 arrays: those are widened by host
        # threads frame by frame); small frames: widened on the device, float64 arrays as ever
        return self.n_w * self.n_h < HOST_WIDEN_MIN_PIXELS or not getattr(self, "_eager_float64", False)

    def begin_block(self, frames, first_index, noise_seed, states, exposure_times):
        """Enqueue ``len(frames)`` consecutive frames, each ONE snapshot ``(unit_time, particles)`` of the same
        number of particles, as one block: the rows of all frames go up back to back, emission / bleaching runs
        frame after frame (the budgets carry over), then binning, rendering and the detector pass run once for
        the whole block (``scb_render_expected_rows_frames``, ``scb_detector_adc_frames``) and every frame is
        downloaded into its own page-locked float32 payload.  Returns one handle per frame for ``finish_frame``
        -- the frames are the ones ``begin_frame`` would have produced (same draws, same accumulator LSBs) -- or
        None when the block must take the frame-by-frame route (ragged snapshots, repeated ids, no payloads)."""
        nb = len(frames)
        sizes = [len(p) for _, p in frames]
        n = sizes[0]
        if nb < 2 or n == 0 or any(m != n for m in sizes):
            return None
        slots_dev = None
        if states is not None:
            slots_dev = []          # per frame: the frames of a movie usually share their ids (one cached device copy)
            for snapshot in frames:
                order, rounds, slots, _ = self._molecule_slots(states.ids, self._ids_of([snapshot]))
                if order is not None or len(rounds) != 1:
                    return None
                slots_dev.append(slots)
        small = self.n_w * self.n_h < HOST_WIDEN_MIN_PIXELS
        if small:
            payloads = [self._host_plane()[0] for _ in range(nb)]       # float64 arrays, widened on the device
        else:
            payloads = [self._lazy_plane() for _ in range(nb)]
            if any(p is None for p in payloads):
                self._lazy_pool["free"].extend(p for p in payloads if p is not None)
                return None
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream(self.device)
            self._frame_stream = ctypes.c_void_p(main.cuda_stream)
            try:
                with _Trace(self, "host_prepare"):
                    return self._begin_block(main, frames, first_index, noise_seed, states, exposure_times, n, slots_dev,
                                             payloads)
            except Exception:
                if not small:
                    self._lazy_pool["free"].extend(payloads)
                raise
            finally:
                self._frame_stream = None

    def _begin_block(self, main, frames, first_index, noise_seed, states, exposure_times, n, slots_dev, payloads):
        cfg = self.configs
        nb = len(frames)
        stream = self._stream()
        focal = cfg.detector_focal_point
        sets = self.__dict__.setdefault("_block_sets", [None, None])
        turn = self._block_turn = (getattr(self, "_block_turn", 0) + 1) % 2
        buf = sets[turn]
        if buf is None or buf["nb"] < nb or buf["n"] < n:
            need = self.lib.scb_render_frames_workspace_bytes(ctypes.byref(self.geom), n, BLOCK_FRAMES)
            if buf is not None and buf["free"] is not None:
                buf["free"].synchronize()
            buf = sets[turn] = dict(
                nb=BLOCK_FRAMES, n=n, free=None,
                rows=torch.empty((BLOCK_FRAMES, n, 5), dtype=torch.float64, device=self.device),
                weight=torch.empty((BLOCK_FRAMES, n), dtype=torch.float64, device=self.device),
                photons=torch.empty((BLOCK_FRAMES, self.n_w, self.n_h), dtype=torch.float32, device=self.device),
                adc=torch.empty((BLOCK_FRAMES, self.n_w, self.n_h), dtype=torch.float32, device=self.device),
                work=torch.empty(int(need) + 256, dtype=torch.uint8, device=self.device),
                det_work=torch.empty(BLOCK_FRAMES * self.lib.scb_detector_workspace_bytes(self.n_w, self.n_h),
                                     dtype=torch.uint8, device=self.device),
                adc64=None, stage=None)
        if buf["free"] is not None:
            main.wait_event(buf["free"])                 # the set's previous block has left the device
        rows = buf["rows"][:nb, :n] if buf["n"] == n else None
        if rows is None:        # fewer particles than the set was made for: a dense view of its memory
            rows = buf["rows"].view(-1)[: nb * n * 5].view(nb, n, 5)
        weight = buf["weight"].view(-1)[: nb * n].view(nb, n)
        photons, adc = buf["photons"][:nb], buf["adc"][:nb]
        keys = []
        all_resident = self.tables.all_resident()
        for k, (unit_time, particles) in enumerate(frames):
            if hasattr(particles, "tensor"):             # base.DeviceRows: the trajectory never left the GPU
                rows[k].copy_(particles.tensor, non_blocking=True)
                if not all_resident:
                    depth = (particles.tensor[:, 0] - float(focal[0])).abs()
                    key = torch.clamp((depth / RESOLUTION).to(torch.int64), max=self.geom.n_depth_keys - 1)
                    key = torch.where(depth < cfg.depth_cutoff + RESOLUTION, key,
                                      torch.full_like(key, self.geom.n_depth_keys))
                    keys.append(torch.unique(key).cpu().numpy())
                continue
            host = None
            if isinstance(particles, numpy.ndarray) and particles.dtype == numpy.float64 and particles.flags.c_contiguous:
                host = torch.from_numpy(particles.view(numpy.ndarray))
                if not host.is_pinned():
                    host = None
            if host is None:                              # through the set's pinned staging area
                if buf["stage"] is None or buf["stage"].shape[1] < n:
                    buf["stage"] = torch.empty((BLOCK_FRAMES, n, 5), dtype=torch.float64, pin_memory=True)
                host = buf["stage"][k, :n]
                numpy.copyto(host.numpy(), numpy.asarray(particles, dtype=numpy.float64))
            rows[k].copy_(host, non_blocking=True)
            if not all_resident:
                keys.append(depth_keys_of(numpy.asarray(particles)[:, 0] - focal[0], cfg.depth_cutoff,
                                          self.geom.n_depth_keys))
        if keys:
            needed = numpy.unique(numpy.concatenate(keys))
            if self.psf_type != _native.PSF_GAUSSIAN and (self.slot_host[needed] < 0).sum() > 128:
                self.ensure_all_tables()
            else:
                self.ensure_tables(needed)
        for k, (unit_time, _) in enumerate(frames):
            self._call(
                "scb_emit_bleach_rows", states.seed if states is not None else 0, n, _native.ptr(rows[k]),
                None if slots_dev is None else _native.ptr(slots_dev[k]), float(unit_time), float(focal[0]),
                ctypes.byref(self.phys), None if states is None else _native.ptr(states.budget),
                _native.ptr(weight[k]), None, stream)
        self._call(
            "scb_render_expected_rows_frames", ctypes.byref(self.geom), n, nb, _native.ptr(rows), _native.ptr(weight),
            _native.ptr(self.sat), _native.ptr(self.box), self.box_type, _native.ptr(self.inv_scale),
            _native.ptr(self.slot_of_key), _native.ptr(photons), _native.F32, _native.ptr(buf["work"]),
            buf["work"].numel(), _native.ptr(self.errors), stream)
        self._call(
            "scb_detector_adc_frames", int(noise_seed), int(first_index), nb, ctypes.byref(self.det), self.n_w, self.n_h,
            _native.F32, _native.ptr(photons), _native.ptr(self.offset), _native.ptr(self.alias),
            0 if self.alias is None else int(self.alias.shape[0]), _native.ptr(adc), _native.ptr(buf["det_work"]),
            buf["det_work"].numel(), stream)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        lazy = payloads[0].dtype == torch.float32
        source = adc
        if not lazy:                # small frames leave the device as float64 (exact widening, one launch per block)
            if buf["adc64"] is None:
                buf["adc64"] = torch.empty((BLOCK_FRAMES, self.n_w, self.n_h), dtype=torch.float64, device=self.device)
            source = buf["adc64"][:nb]
            source.copy_(adc)
        with _Trace(self, "enqueue_d2h"):
            ready = torch.cuda.Event()
            ready.record(main)
            self._copy_stream.wait_event(ready)
            landed = []
            with torch.cuda.stream(self._copy_stream):
                for k, host in enumerate(payloads):
                    host.copy_(source[k], non_blocking=True)
                    event = torch.cuda.Event()
                    event.record(self._copy_stream)
                    landed.append(event)
            buf["free"] = landed[-1]
        errors_host = torch.empty(1, dtype=torch.int32, pin_memory=True)
        errors_host.copy_(self.errors, non_blocking=True)
        self.errors.zero_()
        done = torch.cuda.Event()
        done.record(main)
        return [dict(hosts=[payloads[k]], lazy=lazy, true=None, budget=None, states=states, done=done,
                     errors=errors_host, planes_done=landed[k], tickets=[], exposure_time=exposure_times[k],
                     want_expectation=False) for k in range(nb)]

    def finish_frame(self, pending):
        """Wait for a frame started by ``begin_frame``: ``(adc, expectation or None, true_data,
        budgets dict or None)``.  ``adc`` / ``expectation`` are float64 arrays, or ``image.HostPlane``
        objects (float32 payload, float64 on demand) for large frames begun with ``lazy``."""
        with _Trace(self, "wait_device"):
            pending["done"].synchronize()
            pending["planes_done"].synchronize()
            for ticket in pending["tickets"]:       # float64 planes widened on the host (releases the GIL)
                _native.check(self.lib.scb_host_widen_wait(ticket), "scb_host_widen_wait")
        n_err = int(pending["errors"][0])      # read from pinned memory: no further device sync
        if n_err:
            raise _native.NativeError("{} spots referenced a PSF table that was not built".format(n_err))
        hosts = pending["hosts"]
        if pending["lazy"]:
            from .image import HostPlane
            planes = [HostPlane(self._lazy_view(h), widen=self._float64_wanted) for h in hosts]
        else:
            planes = [h.numpy() for h in hosts]
        adc = planes[0]
        expectation = planes[1] if pending["want_expectation"] else None
        true_data = pending["true"]
        if isinstance(true_data, tuple):
            true_data = self._finish_true_data(true_data[0].numpy().copy(), true_data[1], pending["exposure_time"])
        budgets = None
        if pending["budget"] is not None:
            host = pending["budget"].numpy()
            keep = ~numpy.isnan(host)
            budgets = {int(i): float(b) for i, b in zip(pending["states"].ids[keep], host[keep])}
        return adc, expectation, true_data, budgets

    def abandon_frame(self, pending):
        """Wait until the device and the widening threads are done with a frame that will never be
        handed out (``finish_frame`` was not called): its host arrays may be released afterwards."""
        try:
            pending["done"].synchronize()
            pending["planes_done"].synchronize()
        finally:
            for ticket in pending["tickets"]:
                self.lib.scb_host_widen_wait(ticket)
        if pending.get("lazy"):
            self._lazy_pool["free"].extend(pending["hosts"])      # never handed out: back to the pool

    def __del__(self):
        # staging memory must outlive the widening jobs that read it
        try:
            for tickets in getattr(self, "_stage_tickets", None) or []:
                for ticket in tickets:
                    self.lib.scb_host_widen_wait(ticket)
        except Exception:       # interpreter shutdown: the library may be gone
            pass

    def form_frame(self, snapshots, frame_index, noise_seed, states, exposure_time, want_true_data,
                   want_expectation=True, lazy=False):
        """One frame on the host: ``(adc (Nw, Nh) float64, expectation or None, true_data)``."""
        adc, expectation, true_data, _ = self.finish_frame(self.begin_frame(
            snapshots, frame_index, noise_seed, states, exposure_time, want_true_data, want_expectation, lazy=lazy))
        return adc, expectation, true_data
