"""The C part of the oracle (oracle/overlay_oracle.c) against the numpy oracle."""
import numpy
import pytest

import c_oracle
import epifm_oracle as orc
from conftest import cmos_table, format_inputs, golden, make_configs


@pytest.fixture(scope="module")
def tirf():
    config, configs, params = make_configs("default: {detector: {exposure_time: 0.033}}")
    prof = orc.radial_profile(params, 0.0)
    return config, configs, params, prof


def test_table_and_sat_match_numpy_oracle(tirf):
    _, _, params, prof = tirf
    T = c_oracle.table_from_radial(prof)
    want = orc.PsfTables(params).get(0.0)[1]
    assert abs(T - want).max() / want.max() < 1e-15         # index-domain lerp vs numpy.interp in metres
    S, inv_scale = c_oracle.sat_from_table(T)
    assert inv_scale == 0.5 and S[0].sum() == 0 and S[:, 0].sum() == 0
    assert abs(S[-1, -1] * inv_scale * 1e-18 - 0.9788254597277128) < 1e-12
    # exactness of the integer table: any box sum equals the sum of the quantised samples
    Q = numpy.rint(T / inv_scale).astype(numpy.int64)
    for (a0, a1, b0, b1) in ((0, 1999, 0, 1999), (933, 1000, 950, 1017), (5, 6, 1990, 1999)):
        assert S[a1, b1] - S[a0, b1] - S[a1, b0] + S[a0, b0] == Q[a0:a1, b0:b1].sum()


def test_edges_match_numpy_oracle():
    import ctypes
    rng = numpy.random.RandomState(0)
    lib = c_oracle.lib()
    buf = (ctypes.c_int * 4096)()
    first = ctypes.c_int(0)
    for n_pixel, pl in ((512, 16e-6 / 241), (80, 6.5e-8), (512, 4.444444444444444e-08)):
        for xi in list(rng.uniform(-1.3, 1.3, 200) * n_pixel * pl * 0.5) + [0.0, 5 * n_pixel * pl]:
            n = lib.orc_edges(ctypes.c_double(xi), n_pixel, ctypes.c_double(pl), 1999, ctypes.c_double(1e-9),
                              ctypes.byref(first), buf, 4096)
            i_first, left = orc.overlay_edges(xi, n_pixel, pl, 1999)
            if left is None:
                assert n == 0
            else:
                assert n == len(left) and first.value == i_first and list(buf[:n]) == list(left)


def test_render_forms_agree_on_tirf_c1(tirf):
    config, configs, params, prof = tirf
    g = golden("tirf_c1.npz")
    data = format_inputs(config, g["inputs"])[0][1]
    geom = c_oracle.geometry(params)
    T = c_oracle.table_from_radial(prof)
    S, inv = c_oracle.sat_from_table(T)
    slot = numpy.full(geom.n_depth_keys + 1, -1, dtype=numpy.int32)
    slot[0] = 0
    w = numpy.array([orc.spot_weight(params, orc.emitted(params, d, 0.033), 1.0) for d in data[:, 0]])
    sat = c_oracle.render_sat(geom, data[:, 0], data[:, 1], data[:, 2], w, S[None], [inv], slot)
    brute = c_oracle.render_bruteforce(geom, data[:, 0], data[:, 1], data[:, 2], w, T[None], slot, n_threads=4)
    want = g["photons"]                                        # the live reference's image
    assert abs(sat - want).max() / want.max() < 1e-12
    assert abs(brute - want).max() / want.max() < 1e-12
    assert ((sat > 0) == (want > 0)).all()


def test_detector_port_statistics():
    photons = numpy.full((256, 256), 2.0)
    rn = cmos_table()
    adc = c_oracle.detector_frame(photons, 0.73, 0.01, True, rn[:, 0], rn[:, 1], 0.0, 30000.0, 100.0, 16, seed=3)
    values, p = orc.cmos_readout_pmf(rn)
    gain = 30000.0 / (65536 - 100)
    want = 100 + (0.73 * 2.01 + (values * p).sum()) / gain
    assert abs(adc.mean() - want) < 0.05
    moved = c_oracle.move_points(numpy.zeros((20000, 3)), [1e-8, 2e-8, 0.0], seed=5)
    assert abs(moved[:, 0].std() / 1e-8 - 1) < 0.03 and abs(moved[:, 1].std() / 2e-8 - 1) < 0.03
    assert (moved[:, 2] == 0).all()
