"""The C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

from conftest import ROOT
from scopyon_b200 import _native, build


def header_functions():
    text = open(os.path.join(ROOT, "include", "scopyon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build_library()
    lib = _native.load()
    names = header_functions()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), "{} is declared in the header but not exported".format(name)
    assert sorted(_native.SIGNATURES) == names     # the ctypes prototypes cover the whole header
    assert lib.scb_version() == 100


def test_struct_layouts_match_header():
    # sizes implied by the header's field lists (natural alignment)
    assert ctypes.sizeof(_native.Geometry) == 6 * 4 + 3 * 8 + 3 * 8 + 8      # + box_peak
    assert ctypes.sizeof(_native.Photophysics) == 7 * 8
    assert ctypes.sizeof(_native.Detector) == 4 * 4 + 7 * 8


def test_argument_validation_without_gpu():
    lib = _native.load()
    geom = _native.Geometry(n_w=0, n_h=16, n_radial=1000, n_depth_keys=1002, pixel_length=1e-7, resolution=1e-9,
                            depth_cutoff=1e-6, sat_modulus=65)
    assert lib.scb_render_workspace_bytes(ctypes.byref(geom), 10) == 0
    assert b"image_size" in lib.scb_last_error()
    geom.n_w = 512
    geom.n_h = 512
    assert lib.scb_render_workspace_bytes(ctypes.byref(geom), 100000) > 100000 * 48
    # row sums + scales, and one plain 2000 x 2000 int64 scratch table per key of a build pass
    assert lib.scb_psf_sat_workspace_bytes(1000, 3) == (3 * 1999 + 3) * 8 + 256 + 3 * 2000 * 2000 * 8
    assert lib.scb_psf_sat_slots(1000, 65) == 32 and lib.scb_psf_sat_table_entries(1000, 65) == 65 * 65 * 32 * 32
    assert lib.scb_psf_sat_slots(1000, 1) == 2016 and lib.scb_psf_sat_slots(1000, 100) == 24   # rounded up to a multiple of 4
    rc = lib.scb_psf_radial_build(7, 5e-7, 0.0, 1000, 1, None, None, None)
    assert rc == -2    # SCB_E_NULL, reported before any device work


def test_philox_known_answers():
    """Philox4x32-10 test vectors of Random123 (kat_vectors)."""
    lib = _native.load()
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in cases:
        out = (ctypes.c_uint32 * 4)()
        lib.scb_philox4x32_10((ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_host_widen_is_exact_and_ordered():
    """scb_host_widen_*: the host-thread float32 -> float64 widening of downloaded frames
    (pure host code).  Exact, leaves neighbouring memory alone, tickets complete in order."""
    import numpy
    from scopyon_b200 import _native
    lib = _native.load()
    rng = numpy.random.RandomState(0)
    previous = lib.scb_host_widen_threads(3)
    try:
        for n in (0, 1, 7, 63, 64, 65, 1000, 300 * 301 + 5):
            src = rng.standard_normal(n).astype(numpy.float32)
            if n > 4:
                src[:4] = [numpy.inf, -0.0, numpy.float32(1e-45), numpy.float32(3.4e38)]
            dst = numpy.full(n + 3, -7.0)
            ticket = lib.scb_host_widen_start(src.ctypes.data, dst[1:].ctypes.data, n, None, 0)   # 8-byte aligned only
            assert ticket > 0 and lib.scb_host_widen_wait(ticket) == 0
            assert numpy.array_equal(dst[1:n + 1], src.astype(numpy.float64))
            assert dst[0] == -7.0 and (dst[n + 1:] == -7.0).all()
        src = rng.standard_normal((6, 5000)).astype(numpy.float32)
        dst = numpy.zeros((6, 5000))
        tickets = [lib.scb_host_widen_start(src[k].ctypes.data, dst[k].ctypes.data, 5000, None, 0) for k in range(6)]
        assert tickets == list(range(tickets[0], tickets[0] + 6))
        assert lib.scb_host_widen_wait(tickets[-1]) == 0          # the last one done = all done
        assert numpy.array_equal(dst, src.astype(numpy.float64))
        assert lib.scb_host_widen_wait(tickets[-1] + 1000) != 0 and b"ticket" in lib.scb_last_error()
        assert lib.scb_host_widen_start(None, dst.ctypes.data, 4, None, 0) == -1
    finally:
        lib.scb_host_widen_threads(previous)


def test_host_widen_affinity_and_bandwidth_probe():
    """Worker pinning and the host-bandwidth probe bench.py reports the end-to-end rate against."""
    import numpy
    lib = _native.load()
    cpus = sorted(os.sched_getaffinity(0))[:2]
    assert lib.scb_host_widen_affinity((ctypes.c_int * len(cpus))(*cpus), len(cpus)) == 0
    try:
        src = numpy.arange(100000, dtype=numpy.float32)
        dst = numpy.zeros(100000)
        ticket = lib.scb_host_widen_start(src.ctypes.data, dst.ctypes.data, src.size, None, 0)
        assert ticket > 0 and lib.scb_host_widen_wait(ticket) == 0
        assert numpy.array_equal(dst, src.astype(numpy.float64))
    finally:
        assert lib.scb_host_widen_affinity(None, 0) == 0
    assert lib.scb_host_widen_affinity((ctypes.c_int * 1)(-1), 1) != 0
    for mode in (0, 1):
        rate = ctypes.c_double(0.0)
        assert lib.scb_host_bandwidth(mode, 1 << 20, 2, 2, ctypes.byref(rate)) == 0
        assert rate.value > 1e8          # more than 0.1 GB/s on any machine
    assert lib.scb_host_bandwidth(2, 1 << 20, 2, 2, ctypes.byref(rate)) != 0
