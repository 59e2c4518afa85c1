"""Public facade: the same call signatures as the reference's ``scopyon.base``
(``/root/reference/src/scopyon/base.py:20-294``), with the frame formation handed to
the GPU engine.
"""
import collections.abc
import numbers
import warnings
from logging import getLogger

import numpy

from . import _epifm
from .config import Configuration
from .image import Image

_log = getLogger(__name__)

__all__ = [
    "EnvironSettings", "EPIFMSimulator",
    "form_image", "generate_images", "create_simulator"
    ]


class EnvironSettings:

    def __init__(self, config):
        self.initialize(config)

    def initialize(self, config):
        # multiprocessing.Pool workers of the reference (base.py:20-26); one GPU does the
        # work here, the value is kept for interface compatibility only
        self.processes = config.processes


class ParticleRows(numpy.ndarray):
    """``(N, 5)`` rows ``[depth, x, y, molecule id, p_state]`` as ``__format_data`` returns them --
    a plain float64 array that additionally carries the molecule ids as a contiguous int64
    column (``ids``), and lives in page-locked memory when a GPU is present so the device
    can fetch it with one DMA.  Views and copies are ordinary arrays (``ids`` is dropped)."""

    ids = None

    def __array_finalize__(self, obj):
        self.ids = None


class _PinnedBudget:
    """Bytes of page-locked input rows alive at once (they are freed with their arrays)."""
    limit = 2 << 30
    used = 0

    @classmethod
    def release(cls, nbytes):
        cls.used -= nbytes


def _new_rows(n):
    """Zeroed (n, 5) float64 rows, page-locked when that is possible and affordable."""
    nbytes = n * 5 * 8
    if nbytes and _PinnedBudget.used + nbytes <= _PinnedBudget.limit:
        try:
            import torch
            if torch.cuda.is_available():
                import weakref
                host = torch.zeros((n, 5), dtype=torch.float64, pin_memory=True)
                rows = host.numpy().view(ParticleRows)     # keeps `host` alive through .base
                _PinnedBudget.used += nbytes
                weakref.finalize(host, _PinnedBudget.release, nbytes)
                return rows
        except (ImportError, RuntimeError):
            pass
    return numpy.zeros((n, 5)).view(ParticleRows)


class DeviceRows(object):
    """``(N, 5)`` rows ``[depth, x, y, molecule id, p_state]`` on the GPU (``tensor``) with the molecule ids
    as a host int64 array (``ids``): what ``__format_data`` makes of a ``sampling.DevicePoints`` snapshot."""

    def __init__(self, tensor, ids):
        self.tensor = tensor
        self.ids = ids

    def __len__(self):
        return int(self.tensor.shape[0])


def _project(points, pre):
    """3-D world coordinates -> (depth, x, y) in the camera frame (base.py:76-91)."""
    data = points * pre.scale - numpy.array(pre.origin)
    unit_z = numpy.cross(pre.unit_x, pre.unit_y)
    return numpy.stack((numpy.dot(data, unit_z), numpy.dot(data, pre.unit_x), numpy.dot(data, pre.unit_y)), axis=1)


class EPIFMSimulator(object):

    def __init__(self, config=None, method=None, rng=None):
        """
        Args:
            config (Configuration or str, optional): configuration or a YAML file name.
            method (str, optional): name of the configuration section used
                (defaults to ``config.default``).
            rng (numpy.RandomState, optional): seeds the device random streams.
        """
        if config is None:
            config = Configuration()
        elif isinstance(config, str):
            config = Configuration(filename=config)
        elif not isinstance(config, Configuration):
            raise TypeError("Configuration or str must be given [{}].".format(type(config)))
        if rng is None:
            warnings.warn('A random number generator is not given.')
            rng = numpy.random.RandomState()
        self.__config = config
        # (the reference evaluates `config.default.lower()` here, which cannot work on a
        # sub-tree; fall back to the `method` key like create_simulator does)
        self.__method = method or str(config.get('method', 'default')).lower()
        self.__rng = rng

    def base(self):
        return _epifm._EPIFMSimulator(
            configs=_epifm.EPIFMConfigs(self.__config[self.__method], rng=self.__rng),
            environ=EnvironSettings(self.__config.environ))

    def __format_data(self, inputs):
        """Normalise one point array to ``(N, 5)`` rows ``[depth, x, y, molecule id, p_state]``
        (base.py:61-110)."""
        from .sampling import DevicePoints
        if isinstance(inputs, DevicePoints):
            return self.__format_device(inputs)
        assert isinstance(inputs, numpy.ndarray)
        if inputs.ndim != 2:
            raise ValueError("The given 'inputs' has wrong dimension.")
        pre = self.__config.preprocessing
        n, width = inputs.shape
        data = _new_rows(n)
        if width in (2, 4):     # points on the focal plane
            data[:, 1:3] = inputs[:, :2] * pre.scale
        elif width in (3, 5):
            data[:, 0:3] = _project(inputs[:, :3], pre)
        else:
            raise ValueError("The given 'inputs' has wrong shape.")
        if width in (2, 3):
            data[:, 3] = numpy.arange(n)    # molecule id
            data[:, 4] = 1.0                # photon state
        else:
            data[:, 3:5] = inputs[:, width - 2:]
        data.ids = numpy.ascontiguousarray(data[:, 3]).astype(numpy.int64)
        return data

    def __format_device(self, inputs):
        """``__format_data`` for a snapshot that lives on the GPU: the same row layout, built by tensor
        operations on the device (base.py:61-110)."""
        import torch
        pre = self.__config.preprocessing
        points = inputs.tensor
        n, width = points.shape
        data = torch.zeros((n, 5), dtype=torch.float64, device=points.device)
        if width in (2, 4):
            data[:, 1:3] = points[:, :2] * float(pre.scale)
        elif width in (3, 5):
            as_tensor = lambda v: torch.tensor(numpy.asarray(v, dtype=float), dtype=torch.float64, device=points.device)  # noqa: E731
            shifted = points[:, :3] * float(pre.scale) - as_tensor(pre.origin)
            unit_x, unit_y = numpy.asarray(pre.unit_x, dtype=float), numpy.asarray(pre.unit_y, dtype=float)
            for column, axis in enumerate((numpy.cross(unit_x, unit_y), unit_x, unit_y)):
                data[:, column] = shifted @ as_tensor(axis)
        else:
            raise ValueError("The given 'inputs' has wrong shape.")
        ids = inputs.ids
        if width in (2, 3):
            data[:, 3] = torch.arange(n, dtype=torch.float64, device=points.device)
            data[:, 4] = 1.0
            ids = numpy.arange(n, dtype=numpy.int64)
        else:
            data[:, 3:5] = points[:, width - 2:]
            if ids is None or len(ids) != n:
                ids = points[:, width - 2].cpu().numpy().astype(numpy.int64)
        return DeviceRows(data, ids)

    def __format_inputs(self, inputs):
        from .sampling import DevicePoints
        if isinstance(inputs, (numpy.ndarray, DevicePoints)):
            return ((0.0, self.__format_data(inputs)), )
        if isinstance(inputs, collections.abc.Iterable):
            data = []
            for elem in inputs:
                if not (isinstance(elem, (tuple, list)) and len(elem) == 2
                        and isinstance(elem[0], numbers.Real) and isinstance(elem[1], (numpy.ndarray, DevicePoints))):
                    raise ValueError("The given 'inputs' has wrong type.")
                rows = self.__format_data(elem[1])
                # consecutive snapshots usually show the same molecules in the same order: they then
                # share one id column object, which the engine's per-frame caches recognise by identity
                if data and rows.ids.shape == data[-1][1].ids.shape and numpy.array_equal(rows.ids, data[-1][1].ids):
                    rows.ids = data[-1][1].ids
                data.append((elem[0], rows))
            return data
        raise TypeError(
            "Invalid argument was given [{}]."
            " A ndarray is expected.".format(type(inputs)))

    def form_image(self, inputs, start_time=0.0, exposure_time=None, full_output=False):
        """Form one image.

        Returns:
            Image, or ``(Image, dict)`` when ``full_output``: ``'expectation'`` is the
            expected photoelectron image, ``'true_data'`` maps a molecule id to
            ``[exposure, photon state, X px, Y px, X m, Y m, depth*t, normalization]``.
        """
        data = self.__format_inputs(inputs)
        base = self.base()
        camera, infodict = base.output_frame(
            data, start_time=start_time, exposure_time=exposure_time, rng=self.__rng,
            _full_output=full_output, _planes=not full_output)
        if not full_output:
            return Image(camera)
        img = Image(camera[:, :, 1])
        if full_output:
            infodict.update(dict(expectation=camera[:, :, 0]))
            return img, infodict
        return img

    def generate_images(self, inputs, num_frames, start_time=0.0, exposure_time=None, full_output=False):
        """Generate ``num_frames`` images (a generator), photobleaching carried across frames."""
        data = self.__format_inputs(inputs)
        base = self.base()
        for (camera, infodict) in base.generate_frames(
                data, num_frames, start_time=start_time, exposure_time=exposure_time, rng=self.__rng,
                full_output=full_output):
            if full_output:
                infodict.update(dict(expectation=camera[:, :, 0]))
                yield Image(camera[:, :, 1]), infodict
            else:
                yield Image(camera)     # the bare ADC plane (fast path, no expectation copy)


def create_simulator(config=None, method=None, rng=None):
    """Return a simulator for ``config[method]`` (base.py:203-225)."""
    if method is None:
        method = config.get('method', 'default').lower()
    simulator_type = config[method].type.lower()
    if simulator_type == 'epifm':
        return EPIFMSimulator(config=config, method=method, rng=rng)
    raise ValueError(f"An unknown type [{simulator_type}] was given.")


def form_image(
        inputs, start_time=0.0, exposure_time=None, *,
        method=None, config=None, rng=None, full_output=False):
    """Form one image from points ``(N, 2|3|4|5)`` or ``[(time, points), ...]`` (base.py:227-258)."""
    sim = create_simulator(config, method=method, rng=rng)
    return sim.form_image(inputs, start_time, exposure_time, full_output=full_output)


def generate_images(
        inputs, num_frames, start_time=0.0, exposure_time=None, *,
        method=None, config=None, rng=None, full_output=False):
    """Generate a movie (generator of images), base.py:260-294."""
    sim = create_simulator(config, method=method, rng=rng)
    return sim.generate_images(inputs, num_frames, start_time, exposure_time, full_output=full_output)
