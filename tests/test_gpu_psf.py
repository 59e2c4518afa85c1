"""PSF tables on the GPU against the oracle (C ABI: scb_psf_radial_build, scb_psf_sat_build)."""
import numpy
import pytest

import c_oracle
import epifm_oracle as orc
from conftest import golden, gpu_engine, oracle_tables

pytestmark = pytest.mark.gpu


def test_born_wolf_radial_profile_matches_reference():
    g = golden("radial_profiles.npz")
    _, _, params, engine = gpu_engine()
    keys = [0, 1, 100, 289, 500, 800, 1000, engine.geom.n_depth_keys]   # last = frozen at the cutoff
    engine.ensure_tables(keys)
    got = engine.last_radial.cpu().numpy()
    want = g["born_wolf"]
    # tolerance: CUDA j0/sincos vs cephes/glibc, 100-term sum -> a few ulp of the peak
    assert abs(got - want).max() / want.max() < 1e-13
    rel = abs(got - want) / want.max()
    assert rel.max() < 1e-13 and (got >= 0).all()


def test_gaussian_radial_profile():
    _, _, params, engine = gpu_engine("""
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
""")
    engine.ensure_tables([0, 5, 700])
    assert engine.n_tables == 1 and (engine.slot_host == 0).all()      # depth independent
    got = engine.last_radial.cpu().numpy()[0]
    want = orc.gaussian_radial(orc.radial_grid(1000e-9), 100e-9)
    assert abs(got - want).max() / want.max() < 1e-14


def test_sat_bit_exact_against_c_oracle():
    _, _, params, engine = gpu_engine()
    sats, inv, _ = oracle_tables(params, engine, [0, 800, engine.geom.n_depth_keys])
    got = numpy.stack([engine.tables.plain(k) for k in range(3)])
    # 16 um / 241 = 66.39 nm: 66 x 66 phase blocks of 32 x 32 slots (31 used + 1 past the table per axis)
    assert engine.geom.sat_modulus == 66 and engine.tables.slots == 32
    assert engine.tables.table_entries == 66 * 66 * 32 * 32
    assert numpy.array_equal(engine.inv_scale[:3].cpu().numpy(), inv)
    assert numpy.array_equal(got, sats)                                # int64, bit for bit
    # and the table integral equals the reference's (known answer 0.9788254597277128)
    total = got[0][-1, -1] * inv[0] * 1e-18
    assert abs(total - 0.9788254597277128) < 1e-12
    # slots past the table repeat the last row / column (what the clamped closing edge of a
    # footprint reads): block (pr, pc), slot (ir, ic) holds S[min(ir*M+pr, side)][min(ic*M+pc, side)]
    stored = engine.tables.sat[0].cpu().numpy().reshape(66, 66, 32, 32)
    side = sats[0].shape[0] - 1
    for pr, pc in ((0, 0), (1, 65), (19, 20), (65, 3)):
        a = numpy.minimum(numpy.arange(32) * 66 + pr, side)
        b = numpy.minimum(numpy.arange(32) * 66 + pc, side)
        assert numpy.array_equal(stored[pr, pc], sats[0][a[:, None], b[None, :]])


def test_sat_from_device_profile_close_to_reference_table():
    _, _, params, engine = gpu_engine()
    engine.ensure_tables([100])
    S = engine.tables.plain(0).astype(numpy.float64) * float(engine.inv_scale[0])
    # key 100 is the table at 100 * 1 nm; note int(100e-9 / 1e-9) == 99 in fp64 (_epifm.py:80)
    key, table = orc.PsfTables(params).get(engine.table_depth(100))
    assert key == 100
    want = numpy.zeros_like(S)
    want[1:, 1:] = table.cumsum(0).cumsum(1)
    assert abs(S - want).max() / want.max() < 1e-12


def test_box_table_is_the_pixel_integral_of_each_phase():
    """65 nm pixels (whole number of samples): the box table holds, per phase block, the box sums
    between consecutive slots of the SAT -- what a pixel of an evenly spaced footprint reads."""
    _, _, params, engine = gpu_engine("""
default:
    detector: {type: CMOS, image_size: [64, 64], pixel_length: {value: 6.5e-6, units: m}}
    magnification: 100
""")
    sats, inv, _ = oracle_tables(params, engine, [0, 300])
    assert engine.geom.sat_modulus == 65 and engine.tables.slots == 32 and engine.box is not None
    side = sats[0].shape[0] - 1
    for k in range(2):
        box = engine.box[k].cpu().numpy().reshape(65, 65, 32, 32)
        for pr, pc in ((0, 0), (7, 64), (64, 33), (49, 50)):
            a = numpy.minimum(numpy.arange(32) * 65 + pr, side)
            b = numpy.minimum(numpy.arange(32) * 65 + pc, side)
            s = numpy.zeros((33, 33), dtype=numpy.int64)        # slot -1 is the zero sample before the table
            s[1:, 1:] = sats[k][a[:, None], b[None, :]]
            want = (s[1:, 1:] - s[:-1, 1:]) - (s[1:, :-1] - s[:-1, :-1])
            assert (want >= 0).all()
            assert numpy.array_equal(box[pr, pc], want.astype(numpy.float64))
