"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of the reference's multi-state trajectory generator
(``/root/reference/src/scopyon/sampling2.py``): placement per state, Brownian step with one diffusion
constant per STATE and optional periodic wrap, state transitions through the cumulated one-step
probabilities.  Parity status: **pinned** -- ``generate_points`` / ``move_points`` / ``sample`` are
checked bit for bit against the live reference on the same ``RandomState``
(``tests/test_oracle_vs_reference.py``); the reference's transition step cannot run as written
(``searchsorted(..., side='leff')`` raises ValueError, ``sampling2.py:66``) and is pinned against the
reference executed with that one keyword read as the intended ``'left'``.
"""
import numpy


def generate_points(rng, N, lower, upper, ndim):
    """``sampling2.py:15-31``: rows ``[x.., state, molecule id]``; state ``i`` gets ``N[i]`` points,
    one ``rng.uniform`` call per (state, axis) in that order; degenerate axes sit at ``lower``."""
    ret = numpy.zeros((sum(N), ndim + 2))
    tot = 0
    for i, cnt in enumerate(N):
        for dim in range(ndim):
            if lower[dim] < upper[dim]:
                ret[tot: tot + cnt, dim] = rng.uniform(lower[dim], upper[dim], cnt)
            else:
                ret[tot: tot + cnt, dim] = lower[dim]
        ret[tot: tot + cnt, ndim] = i
        tot += cnt
    ret[:, ndim + 1] = numpy.arange(ret.shape[0])
    return ret


def move_points(rng, points, D, lower, upper, dt, ndim, periodic):
    """``sampling2.py:33-50``: ``x += N(0, sqrt(2 D[state] dt))`` drawn point by point, axis by axis
    (the reference's stream order); then ``(x - lo) % (hi - lo) + lo`` per axis when ``periodic``
    (degenerate axes are reset to ``lower``)."""
    ret = points.copy()
    for i in range(len(ret)):
        scale = numpy.sqrt(2 * D[int(ret[i, ndim])] * dt)
        for dim in range(ndim):
            ret[i, dim] += rng.normal(0.0, scale)
    if periodic:
        for dim in range(ndim):
            if upper[dim] > lower[dim]:
                ret[:, dim] = (ret[:, dim] - lower[dim]) % (upper[dim] - lower[dim]) + lower[dim]
            else:
                ret[:, dim] = lower[dim]
    return ret


def transition_probabilities(transmat, dt):
    """``sampling2.py:56-60``: ``P = 1 - exp(-k dt)`` off the diagonal, the diagonal takes what is
    left of each row; returns the row-wise cumulative sums the draw is searched in."""
    transmat = numpy.asarray(transmat, dtype=float)
    n = transmat.shape[0]
    P = 1 - numpy.exp(-transmat * dt)
    assert (P.sum(axis=1) <= 1.0).all()
    P.ravel()[:: n + 1] = 1.0 - P.sum(axis=1)
    return P.cumsum(axis=1)


def transition_states(rng, points, transmat, dt, ndim):
    """``sampling2.py:52-68`` with the intended ``side='left'`` (``:66`` says ``'leff'``): one uniform per
    point, next state = first column whose cumulated probability reaches it."""
    Pacc = transition_probabilities(transmat, dt)
    ret = points.copy()
    for i in range(ret.shape[0]):
        state = int(ret[i, ndim])
        rnd = rng.uniform(0, 1)
        ret[i, ndim] = numpy.searchsorted(Pacc[state], rnd, side='left')
    return ret


def sample(t, N, lower, upper, D, transmat=None, ndim=3, periodic=False, rng=None):
    """``sampling2.py:70-149`` for already normalised arguments (``N`` list, ``lower``/``upper``/``D``
    arrays): one entry per time point; equal consecutive times repeat the entry."""
    points = generate_points(rng, N, lower, upper, ndim)
    tcurrent = t[0]
    ret = [points.copy()]
    for tnext in t[1:]:
        if tnext > tcurrent:
            dt = tnext - tcurrent
            points = move_points(rng, points, D, lower, upper, dt, ndim, periodic)
            if transmat is not None:
                points = transition_states(rng, points, transmat, dt, ndim)
            tcurrent = tnext
        else:
            assert tnext == tcurrent
        ret.append(points)
    return ret
