"""Host side of the image-formation path.

Mirrors the reference's ``scopyon._epifm`` interface for this path
(``/root/reference/src/scopyon/_epifm.py``): ``EPIFMConfigs`` flattens a
configuration sub-tree into the scalars the kernels need (``:757-996``), and
``_EPIFMSimulator.output_frame / generate_frames`` (``:1017-1049, 1121-1225``) keep
their signatures and return values, but every per-particle and per-pixel loop runs in
``libscopyon_b200.so`` on the GPU.  There is no CPU fallback.

Randomness: the reference consumes a ``numpy.random.RandomState`` draw by draw.  Here
the user's ``rng`` only seeds counter-based Philox streams on the device (one 64-bit
seed per configs object and per frame sequence), so a seeded script is reproducible,
but the streams differ from MT19937's (statistical parity).
"""
import collections
import copy
import json
import math
import os
import warnings
from logging import getLogger

import numpy

from . import _native, constants

_log = getLogger(__name__)

RESOLUTION = 1e-9   # _epifm.py:59-60, depth and radial table pitch
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "catalog_tables.json")
_catalog_cache = None


def catalog_tables():
    """Derived catalog numbers (see tools/make_catalog_tables.py)."""
    global _catalog_cache
    if _catalog_cache is None:
        with open(_DATA) as f:
            _catalog_cache = json.load(f)
    return _catalog_cache


def draw_seed(rng):
    """One 64-bit Philox seed from the user's RandomState."""
    hi, lo = rng.randint(0, 2 ** 32, size=2, dtype=numpy.uint64)
    return (int(hi) << 32) | int(lo)


class PhysicalEffectConfigs:
    """Background, fluorescence and photobleaching settings (``_epifm.py:435-476``)."""

    def __init__(self, config):
        self.set_background(**config.background)
        self.set_fluorescence(**config.fluorescence)
        self.set_photobleaching(**config.photo_bleaching)

    def set_background(self, mean=None, switch=True):
        self.background_switch = switch
        self.background_mean = mean

    def set_fluorescence(self, quantum_yield=None, abs_coefficient=None):
        self.quantum_yield = quantum_yield
        self.abs_coefficient = abs_coefficient

    def set_photobleaching(self, half_life=None, switch=True):
        self.photobleaching_switch = switch
        self.photobleaching_half_life = half_life


class EPIFMConfigs:
    """Flattened microscope settings (``_epifm.py:757-996``)."""

    def __init__(self, config, rng=None):
        if config.type.lower() != 'epifm':
            raise ValueError("An invalid type [{}] was given. 'epifm' is required.".format(config.type))
        if rng is None:
            warnings.warn('A random number generator [rng] is not given.')
            rng = numpy.random.RandomState()

        self.set_fluorophore(**config.fluorophore)
        self.set_shutter(**config.shutter)
        self.set_light_source(**config.light_source)
        self.set_dichroic_mirror(**config.dichroic_mirror)
        self.image_magnification = config.magnification
        self.set_detector(**config.detector)
        self.set_analog_to_digital_converter(rng=rng, **config.analog_to_digital_converter)
        self.set_excitation_filter(**config.excitation_filter)
        self.set_emission_filter(**config.emission_filter)
        self.effects = PhysicalEffectConfigs(config.effects)
        self.radial_cutoff = config.fluorophore.radial_cutoff
        self.depth_cutoff = config.fluorophore.depth_cutoff

    # -- setters keep the reference's keyword names so ``**config.<section>`` expands onto them
    def set_shutter(self, start_time=None, end_time=None, time_open=None, time_lapse=None, switch=True):
        self.shutter_switch = switch
        self.shutter_start_time = start_time
        self.shutter_end_time = end_time

    def set_light_source(self, type=None, wave_length=None, flux_density=None, radius=None, angle=None, switch=True):
        self.source_switch = switch
        self.source_type = type
        self.source_wavelength = wave_length
        self.source_flux_density = flux_density
        self.source_radius = radius
        self.source_angle = angle

    def set_fluorophore(
            self, type=None, wave_length=None, normalization=None, radius=None, radial_width=None,
            min_wave_length=None, max_wave_length=None, radial_cutoff=None, depth_cutoff=None):
        # _epifm.py:822-858.  The emission spectrum only enters through the index of its
        # peak on the 1-nm grid (-> psf_wavelength) and sum(fluoem_norm).
        grid = numpy.arange(min_wave_length, max_wave_length, 1e-9, dtype=float)
        self.fluorophore_type = type
        self.fluorophore_radius = radius
        self.psf_normalization = normalization
        if type == 'Gaussian':
            index_em = int(numpy.abs(grid - wave_length).argmin())
            self.psf_radial_width = radial_width
            self.fluoem_norm_sum = 1.0
        else:
            tables = catalog_tables()
            entry = tables["fluorophore"].get(type)
            if entry is None:
                raise ValueError("An unknown fluorophore type [{}] was given.".format(type))
            ref = tables["wavelength_grid"]
            if not (math.isclose(min_wave_length, ref["min"]) and math.isclose(max_wave_length, ref["max"])):
                raise NotImplementedError(
                    "catalog fluorophores are tabulated on the default wavelength grid "
                    "[{min}, {max}) m only".format(**ref))
            if wave_length is not None:
                warnings.warn('The given wave length [{}] was ignored'.format(wave_length))
            index_em = entry["index_em"]
            self.psf_radial_width = None
            self.fluoem_norm_sum = entry["fluoem_norm_sum"]
        self.psf_wavelength = grid[index_em]

    def set_dichroic_mirror(self, type=None, switch=True):
        self.dichroic_switch = switch
        self._require_filter_off('dichroic_mirror', switch)

    def set_excitation_filter(self, type=None, switch=True):
        self.excitation_switch = switch   # read but never used by the reference (_epifm.py:965-972)

    def set_emission_filter(self, type=None, switch=True):
        self.emission_switch = switch
        self._require_filter_off('emission_filter', switch)

    @staticmethod
    def _require_filter_off(name, switch):
        # In the reference the filter-on branch multiplies a list by a float and raises
        # TypeError (_epifm.py:1310-1313): there is no defined behaviour
        # to reproduce, so the switch must stay off.
        if switch:
            raise NotImplementedError(
                "{}.switch = true is not supported (the reference raises TypeError on this branch)".format(name))

    def set_detector(
            self, type=None, image_size=None, pixel_length=None, exposure_time=None, focal_point=None,
            QE=None, readout_noise=None, dark_count=None, emgain=None, switch=True):
        self.detector_switch = switch
        self.detector_type = type
        self.detector_image_size = image_size
        self.detector_pixel_length = pixel_length
        self.detector_exposure_time = exposure_time
        self.detector_focal_point = focal_point
        self.detector_qeff = QE
        self.detector_readout_noise = readout_noise
        self.detector_dark_count = dark_count
        self.detector_emgain = emgain

    def set_analog_to_digital_converter(self, *, rng, bit=None, offset=None, fullwell=None, type=None, count=None):
        self.ADConverter_bit = bit
        self.ADConverter_fullwell = fullwell
        self.ADConverter_fpn_type = type
        self.ADConverter_fpn_count = count
        self.ADConverter_offset0 = offset
        if type not in _native.FPN_CODES:
            raise ValueError("FPN type [{}] is invalid ['pixel', 'column' or 'none']".format(type))
        if type != 'none' and rng is None:
            raise RuntimeError('A random number generator is required.')
        # The reference draws the offset map here with rng (_epifm.py:943-952); the device
        # draws it from this seed.  Like the reference, a new configs object (one per
        # form_image call, base.py:56-59) gets a new map.
        self.fpn_seed = draw_seed(rng) if type != 'none' else 0

    # -- derived scalars -------------------------------------------------------------
    @property
    def pixel_length(self):
        return self.detector_pixel_length / self.image_magnification   # _epifm.py:1176

    def snells_law(self):
        """(amplitude, penetration depth), ``_epifm.py:1362-1428``."""
        N_0 = self.source_flux_density / (constants.hc / self.source_wavelength)
        sin2 = numpy.sin(self.source_angle) ** 2
        cos2 = numpy.cos(self.source_angle) ** 2
        n_1, n_2 = 1.46, 1.384   # fused silica, cell
        r2 = (n_2 / n_1) ** 2
        if sin2 / r2 < 1:        # epi-illumination
            return N_0, numpy.inf
        A2_x = N_0 * (4 * cos2 * (sin2 - r2) / (r2 ** 2 * cos2 + sin2 - r2))
        A2_y = N_0 * (4 * cos2 / (1 - r2))
        A2_z = N_0 * (4 * cos2 * sin2 / (r2 ** 2 * cos2 + sin2 - r2))
        amplitude = ((A2_x + A2_z) + A2_y) / 2
        depth = self.source_wavelength / (4.0 * numpy.pi * numpy.sqrt(n_1 ** 2 * sin2 - n_2 ** 2))
        return amplitude, depth

    def photophysics(self):
        """The config-only factors of ``get_emit_photons`` (``_epifm.py:1343-1360``)."""
        abs_coeff = self.effects.abs_coefficient
        radius = self.fluorophore_radius
        x_sec = numpy.log(10) * abs_coeff * 0.1 / constants.N_A
        volume = (4.0 / 3.0) * numpy.pi * numpy.power(radius, 3)
        A = (abs_coeff * 0.1 / constants.N_A) * (1.0 / volume) * (2.0 * radius)
        absorb_frac = 1.0 - numpy.power(10.0, -A)
        amplitude0, penetration = self.snells_law()
        budget_scale = 0.0
        if self.effects.photobleaching_switch:
            # get_photon_budget: Exp(scale = half_life/ln2) * N_emit0, N_emit0 at depth 0 for 1 s
            n_emit0 = self.effects.quantum_yield * (amplitude0 * x_sec * 1.0) * absorb_frac
            budget_scale = float(self.effects.photobleaching_half_life / numpy.log(2.0) * n_emit0)
        return _native.Photophysics(
            amplitude0=float(amplitude0), penetration_depth=float(penetration), x_sec=float(x_sec),
            quantum_yield=float(self.effects.quantum_yield), absorb_frac=float(absorb_frac),
            norm_scale=float(self.fluoem_norm_sum * self.psf_normalization), budget_scale=budget_scale)

    def n_radial(self):
        return len(numpy.arange(0.0, self.radial_cutoff, RESOLUTION, dtype=float))   # _epifm.py:91

    def n_depth_keys(self):
        # largest key int(depth / 1 nm) reachable below depth_cutoff + 1 nm (_epifm.py:78-81)
        return int((self.depth_cutoff + RESOLUTION) / RESOLUTION) + 1

    def geometry(self):
        Nw, Nh = self.detector_image_size
        # phase count of the SAT / box-table block layout = pixel pitch in table samples (see include/scopyon_b200.h)
        modulus = int(min(max(round(self.pixel_length / RESOLUTION), 1), 4096))
        g = _native.Geometry(
            n_w=int(Nw), n_h=int(Nh), n_radial=self.n_radial(), n_depth_keys=self.n_depth_keys(),
            sat_modulus=modulus, reserved=0,
            pixel_length=float(self.pixel_length), resolution=RESOLUTION,
            depth_cutoff=float(self.depth_cutoff))
        for i, v in enumerate(self.detector_focal_point):
            g.focal[i] = float(v)
        return g

    def detector_struct(self):
        if self.detector_type not in _native.DETECTOR_CODES:
            raise RuntimeError(
                "Unknown detector type was given [{}]. ".format(self.detector_type)
                + "Use either one of 'CMOS', 'CCD' or 'EMCCD'.")
        return _native.Detector(
            type=_native.DETECTOR_CODES[self.detector_type],
            fpn_type=_native.FPN_CODES[self.ADConverter_fpn_type],
            bit=int(self.ADConverter_bit), background_on=int(bool(self.effects.background_switch)),
            qe=float(self.detector_qeff), background=float(self.effects.background_mean or 0.0),
            readout_noise=float(self.detector_readout_noise or 0.0), emgain=float(self.detector_emgain or 1.0),
            fullwell=float(self.ADConverter_fullwell), adc_offset=float(self.ADConverter_offset0),
            fpn_count=float(self.ADConverter_fpn_count or 0.0))

    def as_oracle_params(self):
        """Plain dict of SI scalars (what the CPU oracle in ``oracle/`` consumes in tests)."""
        return dict(
            image_size=tuple(self.detector_image_size), pixel_length=self.detector_pixel_length,
            magnification=self.image_magnification, focal_point=tuple(self.detector_focal_point),
            psf_type='Gaussian' if self.fluorophore_type == 'Gaussian' else 'BornWolf',
            psf_wavelength=self.psf_wavelength, psf_radial_width=self.psf_radial_width,
            radial_cutoff=self.radial_cutoff, depth_cutoff=self.depth_cutoff,
            psf_normalization=self.psf_normalization, fluoem_norm_sum=self.fluoem_norm_sum,
            source_flux_density=self.source_flux_density, source_wavelength=self.source_wavelength,
            source_angle=self.source_angle, quantum_yield=self.effects.quantum_yield,
            abs_coefficient=self.effects.abs_coefficient, fluorophore_radius=self.fluorophore_radius,
            bleaching_switch=self.effects.photobleaching_switch,
            bleaching_half_life=self.effects.photobleaching_half_life,
            background_switch=self.effects.background_switch, background_mean=self.effects.background_mean,
            detector_type=self.detector_type, QE=self.detector_qeff,
            readout_noise=self.detector_readout_noise, emgain=self.detector_emgain,
            exposure_time=self.detector_exposure_time, adc_bit=self.ADConverter_bit,
            adc_offset=self.ADConverter_offset0, adc_fullwell=self.ADConverter_fullwell,
            fpn_type=self.ADConverter_fpn_type, fpn_count=self.ADConverter_fpn_count,
            shutter_switch=self.shutter_switch, shutter_start_time=self.shutter_start_time,
            shutter_end_time=self.shutter_end_time)


def depth_keys_of(depth_rel, depth_cutoff, n_depth_keys):
    """Vectorised ``PointSpreadingFunction.get`` key rule (``_epifm.py:76-84``); the frozen
    beyond-cutoff table is key ``n_depth_keys``."""
    d = numpy.abs(numpy.asarray(depth_rel, dtype=float))
    keys = numpy.full(d.shape, n_depth_keys, dtype=numpy.int64)
    inside = d < depth_cutoff + RESOLUTION
    keys[inside] = numpy.minimum((d[inside] / RESOLUTION).astype(numpy.int64), n_depth_keys - 1)
    return keys


def frame_windows(times, frame_index, start_time, exposure_time, configs):
    """Snapshots contributing to a frame and their integration times
    (``_epifm.py:1149-1163, 1183-1193``).  Returns ([(snapshot index, unit_time)], t, exposure)."""
    t = start_time + exposure_time * frame_index
    if configs.shutter_switch:
        t = max(t, configs.shutter_start_time)
        exposure_time = max(0.0, min(t + exposure_time, configs.shutter_end_time))
    start_index = numpy.searchsorted(times, t, side='right')
    if start_index != 0:
        start_index -= 1
    stop_index = numpy.searchsorted(times, t + exposure_time, side='left')
    if len(times) > 0 and times[min(start_index, len(times) - 1)] > t:
        warnings.warn("No data input for interval [{}, {}]".format(t, times[start_index]))
    selected = list(range(int(start_index), int(stop_index)))
    windows = []
    for i, k in enumerate(selected):
        current_time = times[k] if i != 0 else t
        next_time = times[selected[i + 1]] if i + 1 < len(selected) else t + exposure_time
        unit_time = next_time - current_time
        if unit_time < 1e-13:   # _epifm.py:1192
            continue
        windows.append((k, float(unit_time)))
    return windows, t, exposure_time


def _payload(plane):
    """The array to copy a finished plane from: itself, or the float32 payload of a ``HostPlane``."""
    return getattr(plane, "f32", plane)


class _EPIFMSimulator:
    """Frame generator with the reference's interface (``_epifm.py:998-1225``)."""

    def __init__(self, configs, environ=None):
        self.configs = configs
        self.environ = environ
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            from .engine import DeviceEngine
            self._engine = DeviceEngine(self.configs)
        return self._engine

    def generate_frames(
            self, input_data, num_frames, start_time=0.0, exposure_time=None,
            rng=None, processes=None, full_output=True):
        """Yield ``(camera, infodict)`` per frame (``_epifm.py:1017-1049``).  The photon
        budgets -- the only state carried from frame to frame -- stay on the device, and
        frames f+1 and f+2 are enqueued before frame f is awaited (two-frame lookahead), so the
        host preparation of the next frames overlaps the device work, the float32 download and
        the host-side widening to float64 of this one."""
        if rng is None:
            _log.info('A random number generator was initialized.')
            rng = numpy.random.RandomState()
        engine = self.engine
        engine._eager_float64 = False       # float32 payloads until the caller asks a frame for its float64 array
        states = None
        if self.configs.effects.photobleaching_switch:
            states = engine.new_budget_state(input_data, draw_seed(rng))
        exposure_time = exposure_time or self.configs.detector_exposure_time
        noise_seed = draw_seed(rng)
        times = numpy.array([t for t, _ in input_data])

        def begin(frame_index):
            windows, t, exposure = frame_windows(times, frame_index, start_time, exposure_time, self.configs)
            _log.info('time: {} sec ({})'.format(t, frame_index))
            snapshots = [(unit_time, input_data[k][1]) for k, unit_time in windows]
            return engine.begin_frame(
                snapshots, frame_index=frame_index, noise_seed=noise_seed, states=states, exposure_time=exposure,
                want_true_data=full_output, want_expectation=full_output, snapshot_states=full_output)

        def finish(pending):
            adc, expectation, true_data, budgets = engine.finish_frame(pending)
            if not full_output:
                return adc, dict(true_data={})      # the bare ADC plane: an array, or a float32 HostPlane
            camera = numpy.empty(adc.shape + (2,), dtype=numpy.float64)   # _epifm.py:1177
            camera[:, :, 0] = _payload(expectation)   # float32 payloads are widened by the assignment (exact)
            camera[:, :, 1] = _payload(adc)
            infodict = dict(true_data=true_data if true_data is not None else {})
            if budgets is not None:
                infodict['fluorescence_states'] = budgets
            return camera, infodict

        def begin_block(first, count):
            """Frames ``first .. first + count - 1`` as one block when each of them is a single snapshot
            (the engine may still decline: ragged snapshots, repeated ids), else None."""
            frames, exposures = [], []
            for frame_index in range(first, first + count):
                windows, t, exposure = frame_windows(times, frame_index, start_time, exposure_time, self.configs)
                if len(windows) != 1:
                    return None
                frames.append((windows[0][1], input_data[windows[0][0]][1]))
                exposures.append(exposure)
            return engine.begin_block(frames, first, noise_seed, states, exposures)

        from . import engine as engine_module
        in_flight = collections.deque()
        block = engine_module.BLOCK_FRAMES if engine.block_route(full_output) else 1
        next_frame = 0
        single_until = 0        # frames below this index go one by one (their block was declined)
        try:
            while next_frame < num_frames or in_flight:
                # blocks: the next one is enqueued while the frames of this one are handed out; single frames:
                # FRAMES_IN_FLIGHT - 1 frames of lookahead
                while next_frame < num_frames:
                    pending = None
                    count = min(block, num_frames - next_frame)
                    if count > 1 and next_frame >= single_until and engine.block_route(full_output):
                        if len(in_flight) >= block:         # next block: as soon as this one's first frame is out
                            break
                        pending = begin_block(next_frame, count)
                        if pending is None:
                            single_until = next_frame + count
                    if pending is None:
                        if len(in_flight) >= engine_module.FRAMES_IN_FLIGHT:
                            break
                        pending = [begin(next_frame)]
                    in_flight.extend(pending)
                    next_frame += len(pending)
                yield finish(in_flight.popleft())
        finally:
            # the caller stopped early (or a frame failed): nothing may still be writing into the
            # host arrays of the frames in flight when they are released
            for pending in in_flight:
                engine.abandon_frame(pending)

    def output_frame(
            self, input_data, frame_index=0, start_time=0.0, exposure_time=None,
            fluorescence_states=None, rng=None, processes=None, _noise_seed=None, _full_output=True,
            _planes=False):
        """One camera frame, ``(camera (Nw, Nh, 2) float64, infodict)`` (``_epifm.py:1121-1225``).

        ``camera[:, :, 0]`` is the expected photoelectron image, ``camera[:, :, 1]`` the ADC
        counts.  ``fluorescence_states`` may be ``None`` (no bleaching), a ``dict`` of
        molecule id -> remaining budget (updated in place like the reference's), or the
        device-resident state ``generate_frames`` creates.  (``_planes=True`` is the facade's
        fast path: ``camera`` is then the bare ``(Nw, Nh)`` ADC image.)
        """
        exposure_time = exposure_time or self.configs.detector_exposure_time
        if rng is None:
            _log.info('A random number generator was initialized.')
            rng = numpy.random.RandomState()
        engine = self.engine
        times = numpy.array([t for t, _ in input_data])
        windows, t, exposure_time = frame_windows(times, frame_index, start_time, exposure_time, self.configs)
        _log.info('time: {} sec ({})'.format(t, frame_index))

        states = fluorescence_states
        from_dict = isinstance(fluorescence_states, dict)
        if from_dict:
            states = engine.new_budget_state(input_data, draw_seed(rng), initial=fluorescence_states)
        if states is not None and not self.configs.effects.photobleaching_switch:
            states = None   # _epifm.py:1295: the budget is only touched when the switch is on

        noise_seed = draw_seed(rng) if _noise_seed is None else _noise_seed
        snapshots = [(unit_time, input_data[k][1]) for k, unit_time in windows]
        adc, expectation, true_data = engine.form_frame(
            snapshots, frame_index=frame_index, noise_seed=noise_seed, states=states,
            exposure_time=exposure_time, want_true_data=_full_output, want_expectation=not _planes, lazy=True)
        if _planes:
            camera = adc
        else:
            camera = numpy.empty(adc.shape + (2,), dtype=numpy.float64)   # _epifm.py:1177
            camera[:, :, 0] = _payload(expectation)
            camera[:, :, 1] = _payload(adc)

        infodict = dict(true_data=true_data if true_data is not None else {})
        if fluorescence_states is not None:
            if from_dict:
                if states is not None:
                    fluorescence_states.update(states.as_dict(only_seen=True))
                infodict['fluorescence_states'] = copy.copy(fluorescence_states)
            elif _full_output:
                infodict['fluorescence_states'] = fluorescence_states.as_dict(only_seen=True)
        return camera, infodict
