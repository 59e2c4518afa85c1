"""Particle kernels (C ABI: scb_place_uniform, scb_diffuse, scb_transition_states,
scb_emit_bleach) and the sampling API built on them."""
import ctypes

import numpy
import pytest
import scipy.stats
import torch

import epifm_oracle as orc
import scopyon_b200
from conftest import gpu_engine
from scopyon_b200 import _native
from scopyon_b200.sampling import DeviceParticles

pytestmark = pytest.mark.gpu


def test_uniform_placement():
    parts = DeviceParticles(200000, 3)
    parts.place_uniform(42, [-1e-5, 0.0, 2e-6], [1e-5, 3e-5, 2e-6])
    xyz = parts.coords.cpu().numpy()
    assert xyz[0].min() >= -1e-5 and xyz[0].max() < 1e-5 and (xyz[2] == 2e-6).all()   # degenerate axis fixed
    assert scipy.stats.kstest((xyz[0] + 1e-5) / 2e-5, "uniform").pvalue > 1e-3
    assert scipy.stats.kstest(xyz[1] / 3e-5, "uniform").pvalue > 1e-3
    assert abs(numpy.corrcoef(xyz[0], xyz[1])[0, 1]) < 0.02


def test_device_philox_matches_known_answer_implementation():
    """The device generator (inline-PTX Philox4x32-10) against the host one that passes the
    Random123 known-answer vectors: placement is lower + u * (upper - lower) with
    u = 53 bits of philox(seed; particle, 0, 'PLAC')."""
    import ctypes
    lib = _native.load()
    seed, n = 0x0123456789ABCDEF, 1000
    parts = DeviceParticles(n, 2)
    parts.place_uniform(seed, [0.0, 0.0], [1.0, 1.0], first_particle=5)
    got = parts.coords.cpu().numpy()
    want = numpy.zeros((2, n))
    key = (ctypes.c_uint32 * 2)(seed & 0xFFFFFFFF, seed >> 32)
    out = (ctypes.c_uint32 * 4)()
    for p in range(n):
        lib.scb_philox4x32_10((ctypes.c_uint32 * 4)(p + 5, 0, 0, 0x504c4143), key, out)
        want[0, p] = (((out[0] << 32) | out[1]) >> 11) * 2.0 ** -53
        want[1, p] = (((out[2] << 32) | out[3]) >> 11) * 2.0 ** -53
    assert numpy.array_equal(got[:2], want)


def test_brownian_msd_and_normality():
    """x += N(0, sqrt(2 D dt)) per axis (sampling.py:116-118): MSD = 2 D dt per axis."""
    n, D, dt = 400000, numpy.array([1e-13, 4e-13, 0.0]), 0.033
    parts = DeviceParticles(n, 3)
    sigma = numpy.sqrt(2 * D * dt)
    parts.step(7, 0, sigma)
    d = parts.coords.cpu().numpy()
    for ax in range(2):
        assert abs(d[ax].var() / sigma[ax] ** 2 - 1) < 5 * numpy.sqrt(2.0 / n)
        assert abs(d[ax].mean()) < 5 * sigma[ax] / numpy.sqrt(n)
        assert scipy.stats.kstest(d[ax] / sigma[ax], "norm").pvalue > 1e-3
    assert (d[2] == 0).all()
    assert abs(numpy.corrcoef(d[0], d[1])[0, 1]) < 0.01
    # ten more steps: variances add
    parts.step(7, 1, sigma, n_steps=10)
    d = parts.coords.cpu().numpy()
    assert abs(d[0].var() / (11 * sigma[0] ** 2) - 1) < 5 * numpy.sqrt(2.0 / n)


def test_trajectory_independent_of_sharding():
    """Keyed (seed; particle, step): shards of particles / of steps reproduce one trajectory."""
    n = 10001
    whole = DeviceParticles(n, 3)
    whole.step(99, 0, [1e-8, 2e-8, 3e-8], n_steps=6)
    want = whole.coords.cpu().numpy()
    # steps split 2 + 4
    split = DeviceParticles(n, 3)
    split.step(99, 0, [1e-8, 2e-8, 3e-8], n_steps=2)
    split.step(99, 2, [1e-8, 2e-8, 3e-8], n_steps=4)
    assert numpy.array_equal(split.coords.cpu().numpy(), want)
    # particles split across two "ranks"
    lo, hi = DeviceParticles(4000, 3), DeviceParticles(n - 4000, 3)
    lo.step(99, 0, [1e-8, 2e-8, 3e-8], n_steps=6)
    hi.step(99, 0, [1e-8, 2e-8, 3e-8], n_steps=6, first_particle=4000)
    assert numpy.array_equal(numpy.concatenate([lo.coords.cpu().numpy(), hi.coords.cpu().numpy()], axis=1), want)
    other = DeviceParticles(n, 3)
    other.step(100, 0, [1e-8, 2e-8, 3e-8], n_steps=6)
    assert not numpy.array_equal(other.coords.cpu().numpy(), want)


def test_sample_inputs_api():
    rng = numpy.random.RandomState(123)
    t = numpy.arange(0, 11 * 0.033, 0.033)
    inputs = scopyon_b200.sample_inputs(t, N=5000, lower=-1e-5, upper=1e-5, ndim=2, D=0.1e-12, rng=rng)
    assert len(inputs) == len(t) and all(p.shape == (5000, 4) for _, p in inputs)
    assert [tt for tt, _ in inputs] == pytest.approx(list(t))
    first, last = inputs[0][1], inputs[-1][1]
    assert (first[:, 2] == numpy.arange(5000)).all() and (first[:, 3] == 1).all()
    assert first[:, :2].min() >= -1e-5 and first[:, :2].max() <= 1e-5
    disp = last[:, :2] - first[:, :2]
    want = 2 * 0.1e-12 * (t[-1] - t[0])
    assert abs(disp.var(axis=0) / want - 1).max() < 5 * numpy.sqrt(2.0 / 5000)
    again = scopyon_b200.sample_inputs(t, N=5000, lower=-1e-5, upper=1e-5, ndim=2, D=0.1e-12,
                                       rng=numpy.random.RandomState(123))
    assert numpy.array_equal(again[-1][1], last)
    # 3-D, per-axis D, explicit start id, repeated time point
    inputs = scopyon_b200.sample_inputs([0.0, 0.1, 0.1, 0.2], N=[30, 20], lower=[0, 0, 0], upper=[1e-5, 1e-5, 1e-6],
                                        D=[1e-13, 0.0, 1e-13], start=7, ndim=3, rng=numpy.random.RandomState(1))
    assert inputs[0][1].shape == (50, 5) and inputs[0][1][0, 3] == 7
    assert numpy.array_equal(inputs[1][1], inputs[2][1])
    assert numpy.array_equal(inputs[0][1][:, 1], inputs[3][1][:, 1])        # D_y = 0
    with pytest.raises(ValueError):
        scopyon_b200.sample_inputs(t, rng=rng)


def test_sample_multistate_periodic_and_transitions():
    rng = numpy.random.RandomState(5)
    t = numpy.arange(0, 2.0, 0.1)
    out = scopyon_b200.sample(t, N=[3000, 3000, 0], lower=0, upper=1e-6, D=[1e-12, 0.0, 0.0], ndim=2,
                              periodic=True, rng=rng)
    assert len(out) == len(t) and out[0].shape == (6000, 4)
    last = out[-1]
    assert last[:, :2].min() >= 0 and last[:, :2].max() < 1e-6              # wrapped into the box
    assert numpy.array_equal(last[3000:, :2], out[0][3000:, :2])            # state 1 does not move
    assert (last[:, 3] == numpy.arange(6000)).all()
    # two-state switching: stationary occupancy k21 / (k12 + k21)
    out = scopyon_b200.sample(t, N=[20000, 0], D=[0.0, 0.0], transmat=[[0.0, 3.0], [1.0, 0.0]], ndim=2,
                              rng=numpy.random.RandomState(6))
    frac1 = (out[-1][:, 2] == 1).mean()
    assert abs(frac1 - 0.75) < 0.02
    step = (out[1][:, 2] == 1).mean()                                       # one step from all-in-0
    assert abs(step - (1 - numpy.exp(-3.0 * 0.1))) < 0.01


def test_emission_and_bleaching_against_oracle():
    yaml = """
default:
    detector: {image_size: [32, 32], exposure_time: 0.033}
    effects: {photo_bleaching: {half_life: {value: 0.2, units: s}}}
"""
    _, configs, params, engine = gpu_engine(yaml)
    rng = numpy.random.RandomState(8)
    n = 50000
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 1.5e-6, n)
    data[:, 3] = numpy.arange(n)[::-1] * 3 + 5
    data[:, 4] = 1.0
    inputs = [(0.0, data)]
    states = engine.new_budget_state(inputs, seed=77)
    dev = lambda a, dt=None: torch.from_numpy(numpy.ascontiguousarray(a)).to(engine.device)
    soa = dev(data[:, [0, 1, 2, 4]].T)
    slots, ids = dev(states.slots_of(data[:, 3])), dev(data[:, 3].astype(numpy.int64))
    w = torch.empty(n, dtype=torch.float64, device=engine.device)

    def emit(unit_time):
        engine._call("scb_emit_bleach", states.seed, n, _native.ptr(soa[0]), _native.ptr(soa[1]), _native.ptr(soa[2]),
                     _native.ptr(soa[3]), _native.ptr(slots), _native.ptr(ids), float(unit_time), 0.0,
                     ctypes.byref(engine.phys), _native.ptr(states.budget), _native.ptr(w), None, engine._stream())
        return w.cpu().numpy()

    emit(0.0)                                                # draws the budgets, emits nothing
    beta, n_emit0 = orc.photon_budget_scale(params)
    initial = states.as_dict()
    draws = numpy.array([initial[int(i)] for i in data[:, 3]]) / (beta * n_emit0)
    assert scipy.stats.kstest(draws, "expon").pvalue > 1e-3   # rng.exponential(scale=beta), _epifm.py:1489
    assert abs(draws.mean() - 1) < 5 / numpy.sqrt(n)
    budgets = dict(initial)
    for step in range(4):
        got = emit(0.033)
        n_emit = numpy.array([orc.emitted(params, d, 0.033) for d in data[:2000, 0]])
        want = numpy.zeros(2000)
        for i in range(2000):
            m = int(data[i, 3])
            b = budgets[m] - n_emit[i]
            p_state = 1.0
            if b <= 0:
                b, p_state = 0, 0.0
            budgets[m] = b
            want[i] = orc.spot_weight(params, n_emit[i], p_state)
        assert numpy.allclose(got[:2000], want, rtol=1e-14, atol=0)
        assert ((got[:2000] == 0) == (want == 0)).all()
    after = states.as_dict()
    assert all(abs(after[m] - budgets[m]) <= 1e-12 * max(1.0, budgets[m]) for m in list(budgets)[:2000]
               if m in {int(i) for i in data[:2000, 3]})
    assert (got == 0).mean() > 0.03                          # a fair share has bleached after 4 x 33 ms at T1/2 = 0.2 s


def test_multistate_sampler_against_the_sampling2_oracle():
    """scopyon_b200.sample (GPU) against oracle/sampling2_oracle.py (the restated sampling2.py with
    `side='left'`): per-state displacement variance 2 D[state] dt per axis and two-sample KS of the
    displacements, periodic wrap into [lower, upper), and the one-step transition counts -- chi-square of
    the GPU's count matrix against the oracle's one-step probabilities, and two-sample against the
    oracle's own draw."""
    import scipy.stats
    import sampling2_oracle as s2
    n_per = 4000
    N = [n_per, n_per, n_per]
    D = numpy.array([1e-12, 0.0, 2.5e-13])
    dt = 0.05
    lower, upper = numpy.zeros(2), numpy.ones(2) * 4e-6
    transmat = numpy.array([[0.0, 2.0, 0.5], [1.0, 0.0, 0.0], [0.3, 4.0, 0.0]])
    t = [0.0, dt]
    # no transitions, no wrap: displacements per state
    got = scopyon_b200.sample(t, N, lower=lower, upper=upper, D=D, ndim=2, rng=numpy.random.RandomState(1))
    want = s2.sample(t, N, lower, upper, D, ndim=2, rng=numpy.random.RandomState(2))
    for state in range(3):
        sel = got[0][:, 2] == state
        d_got = (got[1][sel, :2] - got[0][sel, :2]).ravel()
        d_want = (want[1][want[0][:, 2] == state, :2] - want[0][want[0][:, 2] == state, :2]).ravel()
        sigma = numpy.sqrt(2 * D[state] * dt)
        if sigma == 0:
            assert (d_got == 0).all() and (d_want == 0).all()
            continue
        assert abs(d_got.std() / sigma - 1) < 0.03 and abs(d_got.mean()) < 4 * sigma / numpy.sqrt(len(d_got))
        assert scipy.stats.ks_2samp(d_got, d_want).pvalue > 0.01            # 8000 vs 8000 draws
        assert scipy.stats.kstest(d_got / sigma, "norm").pvalue > 0.01
    # periodic wrap: same rule as the oracle applied to the GPU's own unwrapped step
    wrapped = scopyon_b200.sample(t, N, lower=lower, upper=upper, D=D * 400, ndim=2, periodic=True,
                                  rng=numpy.random.RandomState(1))
    free = scopyon_b200.sample(t, N, lower=lower, upper=upper, D=D * 400, ndim=2, periodic=False,
                               rng=numpy.random.RandomState(1))
    assert (free[1][:, :2] < 0).any() and (free[1][:, :2] > 4e-6).any()      # some molecules did leave the box
    want_wrapped = (free[1][:, :2] - lower) % (upper - lower) + lower           # sampling2.py:47-48
    assert numpy.allclose(wrapped[1][:, :2], want_wrapped, rtol=0, atol=1e-20)
    assert wrapped[1][:, :2].min() >= 0 and wrapped[1][:, :2].max() < 4e-6
    # transitions: one step from known states
    got = scopyon_b200.sample(t, N, lower=lower, upper=upper, D=D, transmat=transmat, ndim=2,
                              rng=numpy.random.RandomState(3))
    want = s2.sample(t, N, lower, upper, D, transmat=transmat, ndim=2, rng=numpy.random.RandomState(4))
    Pacc = s2.transition_probabilities(transmat, dt)
    P = numpy.diff(numpy.concatenate([numpy.zeros((3, 1)), Pacc], axis=1), axis=1)
    for state in range(3):
        counts_got = numpy.bincount(got[1][got[0][:, 2] == state, 2].astype(int), minlength=3)
        counts_want = numpy.bincount(want[1][want[0][:, 2] == state, 2].astype(int), minlength=3)
        keep = P[state] > 0
        assert counts_got[~keep].sum() == 0 and counts_want[~keep].sum() == 0
        assert scipy.stats.chisquare(counts_got[keep], n_per * P[state][keep]).pvalue > 0.001
        assert scipy.stats.chi2_contingency(numpy.stack([counts_got[keep], counts_want[keep]]))[1] > 0.001
