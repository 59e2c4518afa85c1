"""Multi-state trajectory generator on the GPU.

Same interface as the reference's ``scopyon.sampling2.sample``
(``/root/reference/src/scopyon/sampling2.py:70-149``): per-state diffusion constants,
optional periodic box, optional state transitions.  The reference's transition step
calls ``searchsorted(..., side='leff')`` and raises ``ValueError`` (``:66``); the
intended ``side='left'`` is implemented here.
"""
import collections.abc
import ctypes
import warnings
from logging import getLogger

import numpy

from . import _native
from ._epifm import draw_seed
from .sampling import DeviceParticles, _limits

_log = getLogger(__name__)

__all__ = ["sample"]


def sample(t, N, *, lower=None, upper=None, D=None, transmat=None, ndim=3, periodic=False, rng=None):
    """Generate the points: a list (one entry per time point) of arrays of shape
    ``(sum(N), ndim + 2)`` with rows ``[coordinates..., state, molecule id]``."""
    if not isinstance(N, collections.abc.Iterable):
        N = [N]
    N = [int(n) for n in N]
    lower, upper = _limits(lower, upper, ndim)

    if D is None:
        D = numpy.zeros(len(N))
    elif not isinstance(D, collections.abc.Iterable):
        D = numpy.ones(len(N)) * D
    else:
        D = numpy.asarray(D, dtype=float)
        assert len(D) == len(N)

    if transmat is not None:
        transmat = numpy.asarray(transmat, dtype=float)
        n, m = transmat.shape
        assert n == m
        assert (transmat.diagonal() == 0).all()
        assert n == len(N)

    if rng is None:
        warnings.warn('A random number generator [rng] is not given.')
        rng = numpy.random.RandomState()

    total = sum(N)
    seed = draw_seed(rng)
    parts = DeviceParticles(total, ndim)
    torch = parts.torch
    parts.place_uniform(seed, lower[:ndim], upper[:ndim])
    state_host = numpy.repeat(numpy.arange(len(N), dtype=numpy.int32), N)
    state = torch.from_numpy(state_host).to(parts.device)
    ids = numpy.arange(total, dtype=float)

    def snapshot():
        points = numpy.zeros((total, ndim + 2))
        points[:, :ndim] = parts.coords[:ndim].cpu().numpy().T
        points[:, ndim + 0] = state.cpu().numpy()
        points[:, ndim + 1] = ids
        return points

    tcurrent = t[0]
    ret = [snapshot()]
    step_index = 0
    for tnext in t[1:]:
        if tnext > tcurrent:
            dt = tnext - tcurrent
            sigma_state = torch.from_numpy(numpy.sqrt(2 * D * dt)).to(parts.device)   # sampling2.py:41
            parts.step(seed, step_index, None, state=state, sigma_state=sigma_state,
                       periodic=periodic, lower=lower[:ndim], upper=upper[:ndim])
            if transmat is not None:
                # sampling2.py:56-60: P = 1 - exp(-k dt), diagonal = stay probability, row cumsum
                P = 1 - numpy.exp(-transmat * dt)
                assert (P.sum(axis=1) <= 1.0).all()
                P.ravel()[:: len(N) + 1] = 1.0 - P.sum(axis=1)
                pacc = torch.from_numpy(numpy.ascontiguousarray(P.cumsum(axis=1))).to(parts.device)
                _native.check(parts.lib.scb_transition_states(
                    seed, step_index, total, 0, ctypes.c_void_p(state.data_ptr()),
                    ctypes.c_void_p(pacc.data_ptr()), len(N), parts._stream()), "scb_transition_states")
            step_index += 1
            tcurrent = tnext
            ret.append(snapshot())
        else:
            assert tnext == tcurrent
            ret.append(ret[-1])
    return ret
