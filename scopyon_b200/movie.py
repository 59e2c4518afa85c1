"""Device-resident movies, partitioned by frame blocks across GPUs.

The reference makes a movie by materialising the whole trajectory on the host
(``sample_inputs``, ``sampling.py:154-162``) and rendering frames strictly in sequence
(``generate_frames``, ``_epifm.py:1045-1049``).  Here particles live on the GPU: frame
``f`` = [emission + photobleaching at the current positions -> strip-binned PSF render ->
detector/ADC] followed by one Brownian step.  Every random draw is keyed by
(seed; particle|pixel, frame), so rank ``r`` of ``G`` reproduces frames
``[r F/G, (r+1) F/G)`` exactly as a single GPU would: it replays the trajectory and the
photon budgets of the frames before its block with ``scb_replay_frames`` (no rendering)
and then renders its own block.  No collective sits inside the per-frame compute; NCCL is
only used to gather finished frames (``gather_frames``).
"""
import ctypes
import os

import numpy

from . import _native
from ._epifm import EPIFMConfigs, draw_seed
from .engine import DeviceEngine, torch


def frame_block(num_frames, rank, world_size):
    """Contiguous block ``[first, last)`` of frames owned by ``rank``."""
    base, extra = divmod(int(num_frames), int(world_size))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


class DeviceMovie:
    """N molecules diffusing in a box, imaged frame by frame on one GPU.

    Args:
        config (Configuration): scopyon configuration (``config[method]`` is used).
        n_molecules, lower, upper: uniform initial placement in camera coordinates
            ``(x, y, depth)`` [m] (``sample_points``, ``sampling.py:78-83``).
        D: diffusion constant(s) [m^2/s], scalar or per axis ``(x, y, depth)``.
        seed (int): seeds every Philox stream of the movie.
    """

    def __init__(self, config, n_molecules, lower, upper, D, seed, method="default", device=None,
                 precision="f32", exposure_time=None):
        rng = numpy.random.RandomState(seed % (2 ** 32))
        self.configs = EPIFMConfigs(config[method], rng=rng)
        self.engine = DeviceEngine(self.configs, device=device, precision=precision)
        eng = self.engine
        self.n = int(n_molecules)
        self.exposure = float(exposure_time or self.configs.detector_exposure_time)
        self.place_seed = draw_seed(rng)
        self.diffuse_seed = draw_seed(rng)
        self.budget_seed = draw_seed(rng)
        self.noise_seed = draw_seed(rng)
        D = numpy.ones(3) * D if numpy.isscalar(D) else numpy.asarray(D, dtype=float)
        self.sigma_xyd = numpy.sqrt(2 * D * self.exposure)                 # sampling.py:118
        self.lower = numpy.asarray(lower, dtype=float)
        self.upper = numpy.asarray(upper, dtype=float)
        self.is_3d = bool(self.upper[2] > self.lower[2] or self.lower[2] != self.configs.detector_focal_point[0]
                          or self.sigma_xyd[2] > 0)
        # coords rows: x, y, depth (the order scb_diffuse / scb_place_uniform use)
        self.coords = torch.zeros((3, self.n), dtype=torch.float64, device=eng.device)
        self.weight = torch.empty(self.n, dtype=torch.float64, device=eng.device)
        self.bleaching = self.configs.effects.photobleaching_switch
        self.budget = torch.full((self.n,), float("nan"), dtype=torch.float64, device=eng.device) \
            if self.bleaching else None
        self.frame = 0
        self.photons = torch.empty((eng.n_w, eng.n_h), dtype=eng.dtype, device=eng.device)
        self.work = eng._render_workspace(self.n)
        if self.is_3d:
            eng.ensure_all_tables()
        else:
            eng.ensure_tables([0])
        self.reset()

    # ------------------------------------------------------------------ state
    def _p(self, row):
        return ctypes.c_void_p(self.coords[row].data_ptr())

    def reset(self, first_frame=0):
        """Place the molecules and replay to ``first_frame`` (frames before it are not rendered)."""
        with torch.cuda.device(self.engine.device):
            self._reset(first_frame)

    def _reset(self, first_frame):
        eng = self.engine
        _native.check(eng.lib.scb_place_uniform(
            self.place_seed, self.n, 0, self._p(0), self._p(1), self._p(2),
            _native.vec3(self.lower), _native.vec3(self.upper), eng._stream()), "scb_place_uniform")
        if self.budget is not None:
            self.budget.fill_(float("nan"))
        self.frame = 0
        self.__dict__.pop("_order", None)
        if first_frame > 0:
            sigma_dxy = _native.vec3([self.sigma_xyd[2], self.sigma_xyd[0], self.sigma_xyd[1]])
            _native.check(eng.lib.scb_replay_frames(
                self.diffuse_seed, self.budget_seed, 0, int(first_frame), self.n, 0,
                self._p(2), self._p(0), self._p(1), sigma_dxy, self.exposure,
                float(self.configs.detector_focal_point[0]), ctypes.byref(eng.phys),
                _native.ptr(self.budget), eng._stream()), "scb_replay_frames")
            self.frame = int(first_frame)

    # ------------------------------------------------------------------ one frame
    def render_next(self, adc_out, expectation_out=None):
        """Render frame ``self.frame`` into ``adc_out`` (device tensor (Nw, Nh)) and advance."""
        with torch.cuda.device(self.engine.device):
            self._render_next(adc_out, expectation_out)

    def _render_next(self, adc_out, expectation_out=None):
        eng = self.engine
        stream = eng._stream()
        focal = self.configs.detector_focal_point
        _native.check(eng.lib.scb_emit_bleach(
            self.budget_seed, self.n, self._p(2), self._p(0), self._p(1), None, None, None,
            self.exposure, float(focal[0]), ctypes.byref(eng.phys), _native.ptr(self.budget),
            _native.ptr(self.weight), None, stream), "scb_emit_bleach")
        _native.check(eng.lib.scb_render_expected(
            ctypes.byref(eng.geom), self.n, self._p(2), self._p(0), self._p(1), _native.ptr(self.weight),
            _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type, _native.ptr(eng.inv_scale), _native.ptr(eng.slot_of_key),
            _native.ptr(self.photons), eng.elem_type, 0, _native.ptr(self.work), self.work.numel(),
            _native.ptr(eng.errors), stream), "scb_render_expected")
        eng.detect(self.photons, self.frame, self.noise_seed, adc=adc_out, expectation=expectation_out)
        _native.check(eng.lib.scb_diffuse(
            self.diffuse_seed, self.frame, 1, self.n, 0, self._p(0), self._p(1), self._p(2),
            None, None, None, _native.vec3(self.sigma_xyd), None, None, 0, 0, None, None, stream), "scb_diffuse")
        self.frame += 1

    #: frames binned and rendered per launch by ``render_block`` (the binning kernels are
    #: latency bound and the render kernel balances better over more strips: 16 frames per
    #: launch run 1.45x faster than 16 single frames, 32 another 3 % faster, 64 another 1.3 %;
    #: about 150 MB of scratch per frame at C4)
    frames_per_launch = 64
    #: frames per block of ``stream_frames`` (one launch group and one download each; three page-locked
    #: host blocks of that many frames: a shorter block keeps the first frame's latency and the pinned
    #: memory down, and the exported rates sit at the host bound either way)
    export_block_frames = 32

    def render_block(self, out):
        """Fill ``out`` (device tensor (B, Nw, Nh)) with the next B frames.  Bit-identical to B
        calls of ``render_next``; the frames are processed ``frames_per_launch`` at a time:
        one launch advances the molecules through those frames and records positions and
        weights, one pipeline of launches bins and renders all of them, the detector runs per frame."""
        total = out.shape[0]
        k = 0
        with torch.cuda.device(self.engine.device):
            while k < total:
                nf = min(self.frames_per_launch, total - k)
                if nf == 1:
                    self._render_next(out[k])
                else:
                    self._render_frames(out[k: k + nf])
                k += nf
        return out

    # ------------------------------------------------------------------ frames out to the host
    #: export formats of ``stream_frames``: what leaves the device per pixel
    EXPORT_FORMATS = ("f32", "u16", "u8")

    def _export_plan(self, fmt, ring):
        """Buffers of ``stream_frames``: two device blocks (one renders while the other leaves the
        device), their converted copies for the integer formats, ``ring`` page-locked host blocks."""
        eng = self.engine
        key = (fmt, int(ring), self.export_block_frames)
        plans = self.__dict__.setdefault("_export_plans", {})
        plan = plans.get(key)
        if plan is None:
            nb = self.export_block_frames
            shape = (nb, eng.n_w, eng.n_h)
            wire = {"f32": torch.float32, "u16": torch.uint16, "u8": torch.uint8}[fmt]
            plan = plans[key] = dict(
                block=torch.empty((2,) + shape, dtype=torch.float32, device=eng.device),
                wire=None if fmt == "f32" else torch.empty((2,) + shape, dtype=wire, device=eng.device),
                host=torch.empty((ring,) + shape, dtype=wire, pin_memory=True),
                copy_stream=torch.cuda.Stream(device=eng.device))
        return plan

    def stream_frames(self, num_frames, fmt="f32", sink=None, limits=None, ring=3):
        """Render the next ``num_frames`` frames and stream every finished block of
        ``export_block_frames`` frames to page-locked host memory -- the data plane of a sharded movie
        (the reference yields every frame to its caller, ``_epifm.py:1045-1049``; a rank of a
        frame-block partition exports its own block range, nothing is gathered on one GPU).

        The download of block k (copy stream) overlaps the rendering of block k + 1.  ``fmt``:
        ``"f32"`` the frames as computed (exact), ``"u16"`` rounded 16-bit camera counts
        (``scb_frames_to_u16``), ``"u8"`` the 8-bit pictures ``Video.save`` makes
        (``scb_frames_to_8bit``, ``image.py:98-123``) with the fixed ``limits=(cmin, cmax)``
        (default: the ADC's range).  ``sink(first_frame, block)`` is called in frame order with a
        numpy view ``(frames, Nw, Nh)`` of the host ring -- valid during the call only; ``ring`` host
        blocks are in flight.  Returns the number of frames delivered."""
        if fmt not in self.EXPORT_FORMATS:
            raise ValueError("fmt must be one of {}".format(self.EXPORT_FORMATS))
        eng = self.engine
        if eng.dtype != torch.float32:
            raise ValueError("stream_frames exports the float32 pipeline's frames (precision='f32')")
        ring = max(2, int(ring))
        nb = self.export_block_frames
        lib = eng.lib
        with torch.cuda.device(eng.device):
            plan = self._export_plan(fmt, ring)
            main = torch.cuda.current_stream(eng.device)
            copy = plan["copy_stream"]
            stream = ctypes.c_void_p(main.cuda_stream)
            if fmt == "u8":
                cmin, cmax = limits if limits is not None else (0.0, float(2 ** int(self.configs.ADConverter_bit) - 1))
            n_blocks = (int(num_frames) + nb - 1) // nb
            in_flight = {}          # block -> (copied event, first frame, frames, host slot)
            delivered = 0

            def deliver(k):
                nonlocal delivered
                event, first, nf, slot = in_flight.pop(k)
                event.synchronize()
                if sink is not None:
                    sink(first, plan["host"][slot, :nf].numpy())
                delivered += nf

            for k in range(n_blocks):
                nf = min(nb, int(num_frames) - k * nb)
                half, slot = k & 1, k % ring
                if k - ring in in_flight:               # the host slot is handed back by its previous block
                    deliver(k - ring)
                previous = plan.get(("left", half))     # device block `half` has left the device (block k - 2)
                if previous is not None:
                    main.wait_event(previous)
                first = self.frame
                block = plan["block"][half, :nf]
                if nf == 1:
                    self._render_next(block[0])
                else:
                    self._render_frames(block)
                source = block
                if fmt != "f32":
                    source = plan["wire"][half, :nf]
                    n = block.numel()
                    if fmt == "u16":
                        _native.check(lib.scb_frames_to_u16(_native.ptr(block), n, _native.F32, _native.ptr(source),
                                                            stream), "scb_frames_to_u16")
                    else:
                        _native.check(lib.scb_frames_to_8bit(_native.ptr(block), n, _native.F32, None, float(cmin),
                                                             float(cmax), 0.0, 255.0, _native.ptr(source), stream),
                                      "scb_frames_to_8bit")
                rendered = torch.cuda.Event()
                rendered.record(main)
                copy.wait_event(rendered)
                with torch.cuda.stream(copy):
                    plan["host"][slot, :nf].copy_(source, non_blocking=True)
                    copied = torch.cuda.Event()
                    copied.record(copy)
                plan[("left", half)] = copied
                in_flight[k] = (copied, first, nf, slot)
                if k - (ring - 1) in in_flight:         # keep the sink at most ring - 1 blocks behind
                    deliver(k - (ring - 1))
            for k in sorted(in_flight):
                deliver(k)
            self._raise_device_errors(wait=True)
        return delivered

    def _render_frames(self, out):
        eng = self.engine
        nf = out.shape[0]
        stream = eng._stream()
        focal = float(self.configs.detector_focal_point[0])
        cache = getattr(self, "_block", None)
        if cache is None or cache["nf"] < nf:
            need = eng.lib.scb_render_frames_workspace_bytes(ctypes.byref(eng.geom), self.n, nf)
            cache = self._block = dict(
                nf=nf,
                state=torch.empty((4, nf, self.n), dtype=torch.float64, device=eng.device),   # depth, x, y, weight
                photons=torch.empty((nf, eng.n_w, eng.n_h), dtype=eng.dtype, device=eng.device),
                work=torch.empty(int(need) + 256, dtype=torch.uint8, device=eng.device),
                det_work=torch.empty(nf * eng.lib.scb_detector_workspace_bytes(eng.n_w, eng.n_h), dtype=torch.uint8,
                                     device=eng.device))
        state, photons, work = cache["state"], cache["photons"], cache["work"]
        sigma_dxy = _native.vec3([self.sigma_xyd[2], self.sigma_xyd[0], self.sigma_xyd[1]])
        _native.check(eng.lib.scb_movie_frames(
            self.diffuse_seed, self.budget_seed, self.frame, nf, self.n, 0,
            self._p(2), self._p(0), self._p(1), sigma_dxy, self.exposure, focal, ctypes.byref(eng.phys),
            _native.ptr(self.budget), _native.ptr(state[0]), _native.ptr(state[1]), _native.ptr(state[2]),
            _native.ptr(state[3]), stream), "scb_movie_frames")
        self._raise_device_errors(wait=False)
        # consecutive blocks of the same shape: the binning of this block runs in one pass over the list plan the
        # previous block left in the workspace (bit 0), and leaves one for the next (bit 1); the images are the same
        plan_key = (nf, work.data_ptr(), self.frame)
        plan_mode = 0
        if self.plan_blocks:
            plan_mode = 2 | (1 if cache.get("plan") == plan_key else 0)
        _native.check(eng.lib.scb_render_expected_frames_planned(
            ctypes.byref(eng.geom), self.n, nf, _native.ptr(self._visiting_order(state)),
            _native.ptr(state[0]), _native.ptr(state[1]), _native.ptr(state[2]),
            _native.ptr(state[3]), _native.ptr(eng.sat), _native.ptr(eng.box), eng.box_type,
            _native.ptr(eng.inv_scale), _native.ptr(eng.slot_of_key), _native.ptr(photons), eng.elem_type,
            _native.ptr(work), work.numel(), _native.ptr(eng.errors), plan_mode, stream),
            "scb_render_expected_frames_planned")
        cache["plan"] = (nf, work.data_ptr(), self.frame + nf) if plan_mode & 2 else None
        if plan_mode & 1:
            # a plan that fell short by more than the overflow list holds is counted in the engine's error word:
            # fetched without waiting, looked at when the next block is enqueued (or by check_errors())
            seen = self.__dict__.setdefault("_errors_seen", dict(
                host=torch.zeros(1, dtype=torch.int32).pin_memory(), event=torch.cuda.Event()))
            seen["host"].copy_(eng.errors.reshape(-1)[:1], non_blocking=True)
            seen["event"].record(torch.cuda.current_stream(eng.device))
            seen["pending"] = True
        batched = (eng.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous()
                   and (eng.n_w * eng.n_h) % 4 == 0
                   and (self.configs.ADConverter_fpn_type != 'column' or eng.n_h % 4 == 0))
        if batched:
            offset = eng.offset
            _native.check(eng.lib.scb_detector_adc_frames(
                int(self.noise_seed), int(self.frame), nf, ctypes.byref(eng.det), eng.n_w, eng.n_h, _native.F32,
                _native.ptr(photons), _native.ptr(offset), _native.ptr(eng.alias),
                0 if eng.alias is None else int(eng.alias.shape[0]), _native.ptr(out),
                _native.ptr(cache["det_work"]), cache["det_work"].numel(), stream), "scb_detector_adc_frames")
        else:
            for f in range(nf):
                eng.detect(photons[f], self.frame + f, self.noise_seed, adc=out[f])
        self.weight = state[3, nf - 1]
        self.frame += nf

    def _raise_device_errors(self, wait):
        seen = self.__dict__.get("_errors_seen")
        if not seen or not seen.get("pending"):
            return
        if wait:
            seen["event"].synchronize()
        elif not seen["event"].query():
            return
        seen["pending"] = False
        if int(seen["host"][0]) > 0:
            raise RuntimeError("scopyon_b200: {} device-side errors while rendering a planned block (a list plan fell "
                               "short by more than the overflow list holds): render again with plan_blocks = False"
                               .format(int(seen["host"][0])))

    def check_errors(self):
        """Wait for the blocks enqueued so far and raise if the device counted an error in one of them."""
        with torch.cuda.device(self.engine.device):
            self._raise_device_errors(wait=True)

    #: blocks rendered with one visiting order before it is refreshed (molecules move about a pixel per frame)
    order_refresh_blocks = 8
    #: consecutive blocks bin their spots in one pass over a list plan taken from the previous block's census
    #: (``scb_render_expected_frames_planned``; SCOPYON_B200_PLAN=0 turns it off: count, scan, fill)
    plan_blocks = os.environ.get("SCOPYON_B200_PLAN", "1") != "0"

    def _visiting_order(self, state):
        """Particle indices sorted by the coarse screen cell (32 x 32 pixels, row-major) of the first frame
        of the block: neighbours in the spot list are then neighbours on the screen, which lets the census
        of the binning add a warp's overlaps with a tenth of the atomics
        (``scb_render_expected_frames_ordered``).  The images do not depend on it."""
        cache = self.__dict__.setdefault("_order", dict(age=self.order_refresh_blocks, order=None))
        if cache["order"] is None or cache["age"] >= self.order_refresh_blocks:
            eng = self.engine
            pl = float(self.configs.pixel_length)
            focal = self.configs.detector_focal_point
            row = ((state[1, 0] - float(focal[1])) / pl + eng.n_w * 0.5).clamp_(0, eng.n_w - 1).to(torch.int32) >> 5
            col = ((state[2, 0] - float(focal[2])) / pl + eng.n_h * 0.5).clamp_(0, eng.n_h - 1).to(torch.int32) >> 5
            cache["order"] = torch.argsort(row * ((eng.n_h + 31) >> 5) + col).to(torch.int32)
            cache["age"] = 0
        cache["age"] += 1
        return cache["order"]

    def positions(self):
        """Current ``(N, 5)`` rows ``[depth, x, y, id, p_state]`` on the host (for parity checks)."""
        xyd = self.coords.cpu().numpy()
        data = numpy.zeros((self.n, 5))
        data[:, 0], data[:, 1], data[:, 2] = xyd[2], xyd[0], xyd[1]
        data[:, 3] = numpy.arange(self.n)
        data[:, 4] = 1.0
        return data


class NpyFrameWriter(object):
    """Streaming ``.npy`` writer for a frame stack of known length: the header for the full
    ``(num_frames, Nw, Nh)`` array is written first, blocks of frames are appended as they arrive
    (``sink`` of ``DeviceMovie.stream_frames``), so a movie far larger than host memory goes to
    disk block by block.  The file is what ``numpy.save`` of the whole stack would produce --
    the ``.npy`` route of ``Image.save`` / ``Video.save`` (``image.py:125-140, 240-277``)."""

    def __init__(self, filename, num_frames, frame_shape, dtype):
        self.shape = (int(num_frames),) + tuple(int(v) for v in frame_shape)
        self.dtype = numpy.dtype(dtype)
        self.written = 0
        self.file = open(str(filename), "wb")
        header = {"descr": numpy.lib.format.dtype_to_descr(self.dtype), "fortran_order": False, "shape": self.shape}
        numpy.lib.format.write_array_header_1_0(self.file, header)

    def __call__(self, first_frame, block):
        """Append ``block`` (frames, Nw, Nh); blocks must arrive in frame order."""
        if block.dtype != self.dtype or tuple(block.shape[1:]) != self.shape[1:]:
            raise ValueError("block of {} {} does not fit a {} {} stack".format(block.dtype, block.shape, self.dtype, self.shape))
        if self.written + block.shape[0] > self.shape[0]:
            raise ValueError("more frames than the header announced")
        numpy.ascontiguousarray(block).tofile(self.file)
        self.written += block.shape[0]

    def close(self):
        if self.file is not None:
            self.file.close()
            self.file = None
            if self.written != self.shape[0]:
                raise ValueError("{} of {} frames were written".format(self.written, self.shape[0]))

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.close()
        elif self.file is not None:
            self.file.close()
            self.file = None


def save_movie(movie, filename, num_frames, fmt="f32", limits=None, as_float64=False):
    """Render the next ``num_frames`` frames of ``movie`` (a ``DeviceMovie``) straight into a ``.npy``
    stack on disk: float32 frames (or float64, widened on the host, with ``as_float64``), rounded
    uint16 camera counts, or the uint8 pictures of ``Video.save(filename.npy, ...)`` with fixed
    ``limits``.  Returns the number of frames written."""
    eng = movie.engine
    dtype = {"f32": numpy.float64 if as_float64 else numpy.float32, "u16": numpy.uint16, "u8": numpy.uint8}[fmt]
    with NpyFrameWriter(filename, num_frames, (eng.n_w, eng.n_h), dtype) as writer:
        if fmt == "f32" and as_float64:
            sink = lambda first, block: writer(first, block.astype(numpy.float64))   # noqa: E731 -- exact
        else:
            sink = writer
        return movie.stream_frames(num_frames, fmt=fmt, sink=sink, limits=limits)


def gather_frames(local_frames, num_frames, group=None, dst=None):
    """Assemble the frame stack from per-rank blocks (``frame_block`` partition).

    ``local_frames``: tensor ``(frames of this rank, Nw, Nh)``.  Uses
    ``torch.distributed`` (NCCL over NVLink for device tensors, gloo for host tensors):
    ``all_gather`` of equally padded blocks, or ``gather`` to ``dst``.  Returns the full
    ``(num_frames, Nw, Nh)`` tensor (on ``dst`` only when given, else on every rank).
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    longest = max(frame_block(num_frames, r, world)[1] - frame_block(num_frames, r, world)[0] for r in range(world))
    shape = (longest,) + tuple(local_frames.shape[1:])
    padded = local_frames.new_zeros(shape)
    padded[: local_frames.shape[0]].copy_(local_frames)
    if dst is None:
        blocks = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(blocks, padded, group=group)
    else:
        blocks = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, blocks, dst=dst, group=group)
        if rank != dst:
            return None
    parts = []
    for r in range(world):
        first, last = frame_block(num_frames, r, world)
        parts.append(blocks[r][: last - first])
    return torch.cat(parts, dim=0)


def assemble_channels(channel, group=None):
    """Two-colour assembly (``examples/twocolor.py:14-16``, ``image.py:38-68``): every rank
    renders one channel; all ranks receive the ``(Nw, Nh, world)`` stack."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    planes = [torch.empty_like(channel) for _ in range(world)]
    dist.all_gather(planes, channel.contiguous(), group=group)
    return torch.stack(planes, dim=-1)


def frames_to_8bit(frames, cmin=None, cmax=None, low=None, high=None):
    """8-bit copy of a device frame stack ``(F, Nw, Nh)`` (or one frame), scaled like
    ``Image.as_8bit`` with the limits ``Video.save`` uses: one common ``(cmin, cmax)`` --
    by default the extrema over the whole stack (``image.py:98-123, 261-264``).  The result
    stays on the device (a quarter of the fp32 bytes to download)."""
    if not frames.is_cuda or frames.dtype not in (torch.float32, torch.float64):
        raise TypeError("frames_to_8bit expects a float32/float64 CUDA tensor")
    frames = frames.contiguous()
    lib = _native.load()
    stream = ctypes.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
    elem = _native.F32 if frames.dtype == torch.float32 else _native.F64
    n = frames.numel()
    limits = None
    if cmin is None or cmax is None:
        limits = torch.empty(2, dtype=torch.float64, device=frames.device)
        scratch = torch.empty(2, dtype=torch.int64, device=frames.device)
        _native.check(lib.scb_frames_minmax(_native.ptr(frames), n, elem, _native.ptr(limits), _native.ptr(scratch),
                                            stream), "scb_frames_minmax")
        if cmin is not None:
            limits[0] = float(cmin)
        if cmax is not None:
            limits[1] = float(cmax)
    out = torch.empty(frames.shape, dtype=torch.uint8, device=frames.device)
    _native.check(lib.scb_frames_to_8bit(
        _native.ptr(frames), n, elem, _native.ptr(limits), float(cmin or 0.0), float(cmax or 0.0),
        float(0 if low is None else low), float(255 if high is None else high), _native.ptr(out), stream),
        "scb_frames_to_8bit")
    return out
