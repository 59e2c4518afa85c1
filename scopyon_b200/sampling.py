"""Input generator: uniform placement + Brownian trajectories, computed on the GPU.

Same interface as the reference's ``scopyon.sampling``
(``/root/reference/src/scopyon/sampling.py:14-162``).  The reference draws one
``rng.normal`` per coordinate in a Python double loop (``:116-118``); here a step of
all particles is one kernel launch, and the whole trajectory is a pure function of
(seed, particle index, step index) through the counter-based Philox generator.
"""
import collections.abc
import ctypes
import numbers
import warnings
from logging import getLogger

import numpy

from . import _native
from ._epifm import draw_seed

_log = getLogger(__name__)

__all__ = ["sample_inputs", "DevicePoints"]


class DevicePoints(object):
    """One snapshot of a trajectory that stayed on the GPU: ``tensor`` is the ``(N, ndim + 2)`` float64
    CUDA tensor of the rows ``[coordinates..., molecule id, p_state]`` that ``sample_inputs`` otherwise
    returns as a numpy array, ``ids`` the molecule ids as a host int64 array (one object shared by all
    snapshots of a trajectory).  ``form_image`` / ``generate_images`` accept it wherever they accept the
    array: nothing is uploaded per frame.  (An extension: the reference keeps trajectories on the host.)"""

    def __init__(self, tensor, ids):
        assert tensor.dim() == 2 and tensor.is_cuda
        self.tensor = tensor
        self.ids = ids

    ndim = 2
    shape = property(lambda self: tuple(self.tensor.shape))

    def __len__(self):
        return int(self.tensor.shape[0])

    def cpu(self):
        """The snapshot as the numpy array the host path returns."""
        return self.tensor.cpu().numpy()

_CHUNK_BYTES = 1 << 30   # device staging for trajectories, copied back chunk by chunk


def _limits(lower, upper, ndim):
    """Normalise box limits (``sampling.py:42-55``)."""
    lower = (numpy.ones(ndim) * lower if isinstance(lower, numbers.Number)
             else numpy.array(lower, dtype=float) if lower is not None
             else numpy.zeros(ndim))
    upper = (numpy.ones(ndim) * upper if isinstance(upper, numbers.Number)
             else numpy.array(upper, dtype=float) if upper is not None
             else numpy.ones(ndim))
    if len(lower) < ndim or len(upper) < ndim:
        raise ValueError(
            "The wrong size of limits was given [(lower={}, upper={}) != {}].".format(len(lower), len(upper), ndim))
    for dim in range(ndim):
        if lower[dim] > upper[dim]:
            lower[dim], upper[dim] = upper[dim], lower[dim]
    return lower, upper


def _pad3(values, fill=0.0):
    out = [fill, fill, fill]
    for i, v in enumerate(list(values)[:3]):
        out[i] = float(v)
    return _native.vec3(out)


class DeviceParticles:
    """SoA coordinates of N particles on the device (up to 3 axes; axes beyond 3, which
    the reference permits, never move and stay on the host)."""

    def __init__(self, n, ndim, device=None):
        import torch
        from .engine import require_cuda
        require_cuda()
        self.torch = torch
        self.lib = _native.load()
        self.n, self.ndim = int(n), int(ndim)
        if not 1 <= self.ndim <= 3:
            raise ValueError("ndim must be 1, 2 or 3 on the device path [{}]".format(ndim))
        self.device = torch.device(device if device is not None else "cuda:{}".format(torch.cuda.current_device()))
        self.coords = torch.zeros((3, self.n), dtype=torch.float64, device=self.device)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _axis(self, tensor, d):
        return ctypes.c_void_p(tensor[d].data_ptr()) if d < self.ndim else None

    def place_uniform(self, seed, lower, upper, first_particle=0):
        """``sample_points`` placement (``sampling.py:78-83``)."""
        _native.check(self.lib.scb_place_uniform(
            seed, self.n, first_particle, self._axis(self.coords, 0), self._axis(self.coords, 1),
            self._axis(self.coords, 2), _pad3(lower), _pad3(upper), self._stream()), "scb_place_uniform")

    def step(self, seed, step_index, sigma, out=None, n_steps=1, state=None, sigma_state=None,
             periodic=False, lower=None, upper=None, first_particle=0):
        """``move_points`` (``sampling.py:88-119``): coords += N(0, sigma[d]); writes into
        ``out`` (a (3, N) tensor) when given, else in place."""
        dst = self.coords if out is None else out
        _native.check(self.lib.scb_diffuse(
            seed, step_index, n_steps, self.n, first_particle,
            self._axis(self.coords, 0), self._axis(self.coords, 1), self._axis(self.coords, 2),
            self._axis(dst, 0), self._axis(dst, 1), self._axis(dst, 2),
            None if sigma is None else _pad3(sigma),
            None if state is None else ctypes.c_void_p(state.data_ptr()),
            None if sigma_state is None else ctypes.c_void_p(sigma_state.data_ptr()),
            0 if sigma_state is None else int(sigma_state.numel()),
            int(bool(periodic)), None if lower is None else _pad3(lower),
            None if upper is None else _pad3(upper), self._stream()), "scb_diffuse")


def sample_inputs(t, *, N=None, conc=None, lower=None, upper=None, D=None, start=0, ndim=3, rng=None,
                  device=False):
    """Generate the input data: a list of ``(time, points)`` with ``points`` of shape
    ``(N, ndim + 2)``, rows ``[coordinates..., molecule id, p_state = 1]``.

    Args and return value as in the reference (``sampling.py:121-162``).  ``device=True`` (an extension)
    leaves the snapshots on the GPU as ``DevicePoints`` -- the same numbers, never copied to the host --
    which ``generate_images`` / ``form_image`` take in place of the arrays.
    """
    if rng is None:
        warnings.warn('A random number generator [rng] is not given.')
        rng = numpy.random.RandomState()
    if N is None and conc is None:
        raise ValueError('Either one of N or conc must be given.')

    t = sorted(t)
    lower, upper = _limits(lower, upper, ndim)
    if N is None:   # sampling.py:59-66
        lengths = upper - lower
        size = numpy.prod(lengths[lengths != 0])
        if isinstance(conc, collections.abc.Iterable):
            N_list = [rng.poisson(size * conc_) for conc_ in conc]
        else:
            N_list = [rng.poisson(size * conc)]
    elif not isinstance(N, collections.abc.Iterable):
        N_list = [N]
    else:
        N_list = N
    N = int(sum(N_list))
    if N <= 0:
        empty = numpy.array([])
        return [(tk, empty.copy()) for tk in t]
    if device and len(lower) > 3:
        raise ValueError("device=True supports up to three axes")

    maxdim = len(lower)
    if D is None:
        D = numpy.zeros(ndim)
    elif not isinstance(D, collections.abc.Iterable):
        D = numpy.ones(ndim) * D
    else:
        D = numpy.asarray(D, dtype=float)
        assert len(D) == ndim

    seed = draw_seed(rng)
    dev_dim = min(maxdim, 3)
    parts = DeviceParticles(N, dev_dim)
    parts.place_uniform(seed, lower[:dev_dim], upper[:dev_dim])
    move_dim = min(ndim, dev_dim)

    torch = parts.torch
    n_times = len(t)
    per_step = 3 * N * 8
    chunk = max(1, min(n_times, _CHUNK_BYTES // per_step))
    template = numpy.zeros((N, maxdim + 2))
    for dim in range(dev_dim, maxdim):          # static extra axes (host only)
        template[:, dim] = rng.uniform(lower[dim], upper[dim], N) if lower[dim] < upper[dim] else lower[dim]
    template[:, maxdim + 0] = numpy.arange(start, start + N)   # molecule id
    template[:, maxdim + 1] = 1.0                              # photon state

    inputs = []
    tcurrent = t[0]
    step_index = 0
    k = 0
    if device:
        ids = numpy.arange(start, start + N, dtype=numpy.int64)
        tail = torch.from_numpy(template[:, maxdim:]).to(parts.device)          # (N, 2): id, p_state
        for i in range(n_times):
            tnext = t[i]
            if tnext > tcurrent:
                sigma = [0.0, 0.0, 0.0]
                for dim in range(move_dim):
                    sigma[dim] = float(numpy.sqrt(2 * D[dim] * (tnext - tcurrent)))
                parts.step(seed, step_index, sigma)
                step_index += 1
                tcurrent = tnext
            rows = torch.empty((N, maxdim + 2), dtype=torch.float64, device=parts.device)
            rows[:, :dev_dim] = parts.coords[:dev_dim].t()
            rows[:, maxdim:] = tail
            inputs.append((float(t[i] if t[i] > t[0] else t[0]), DevicePoints(rows, ids)))
        return inputs
    while k < n_times:
        m = min(chunk, n_times - k)
        stage = torch.empty((m, 3, N), dtype=torch.float64, device=parts.device)
        for i in range(m):
            tnext = t[k + i]
            if tnext > tcurrent:
                sigma = [0.0, 0.0, 0.0]
                for dim in range(move_dim):
                    sigma[dim] = float(numpy.sqrt(2 * D[dim] * (tnext - tcurrent)))
                parts.step(seed, step_index, sigma)
                step_index += 1
                tcurrent = tnext
            stage[i].copy_(parts.coords)
        host = stage.cpu().numpy()
        for i in range(m):
            points = template.copy()
            points[:, :dev_dim] = host[i, :dev_dim].T
            inputs.append((t[k + i] if t[k + i] > t[0] else t[0], points))
        k += m
    # the reference labels every entry with the running 'tcurrent'
    return [(float(tk), p) for tk, p in inputs]
