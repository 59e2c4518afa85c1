"""Expected-image rendering on the GPU (C ABI: scb_emit_bleach, scb_render_expected)
against the oracle and the golden vectors from the live reference."""
import numpy
import pytest
import torch

import c_oracle
import epifm_oracle as orc
from conftest import format_inputs, golden, gpu_engine, oracle_tables

pytestmark = pytest.mark.gpu

REL = 1e-9   # max|diff| / max(ref); north_star asks 1e-5, the fp64 SAT path does far better


def rel_err(got, want):
    return abs(got - want).max() / want.max()


def render(engine, data, unit_time=0.033, dtype=torch.float64):
    out = torch.empty((engine.n_w, engine.n_h), dtype=dtype, device=engine.device)
    img, _ = engine.render_expected([(unit_time, data)], out=out)
    torch.cuda.synchronize()
    assert int(engine.errors.item()) == 0
    return img.cpu().numpy()


def test_tirf_c1_expectation_matches_reference(known_answers):
    g = golden("tirf_c1.npz")
    config, _, params, engine = gpu_engine("default: {detector: {exposure_time: 0.033}}")
    data = format_inputs(config, g["inputs"])[0][1]
    got = render(engine, data)
    assert rel_err(got, g["photons"]) < REL
    assert abs(got.sum() - known_answers["tirf_c1"]["photons_sum"]) / got.sum() < 1e-12
    assert ((got > 0) == (g["photons"] > 0)).all()          # identical pixel footprints
    got32 = render(engine, data, dtype=torch.float32)
    assert rel_err(got32.astype(numpy.float64), g["photons"]) < 2e-7


FIXED = 1e-12   # fixed-point accumulators: LSB <= 2^-62 * 2 * n_spots * max weight per term


def test_against_c_oracle_sat_form_and_order_independence():
    """Same tables in, same index arithmetic -> the C oracle's image to the accumulator LSB,
    and identical bits for any order of the spots (integer accumulation is associative)."""
    g = golden("border_depth_case.npz")
    config, configs, params, engine = gpu_engine("default: {detector: {image_size: [96, 80], exposure_time: 0.033}}")
    data = g["formatted"]
    keys = numpy.unique(engine_keys(engine, configs, data))
    sats, inv, slot = oracle_tables(params, engine, keys)
    n_emit = numpy.array([orc.emitted(params, d, 0.033) for d in data[:, 0]])
    weight = numpy.array([orc.spot_weight(params, e, 1.0) for e in n_emit])
    want = c_oracle.render_sat(c_oracle.geometry(params), data[:, 0], data[:, 1], data[:, 2], weight, sats, inv, slot)
    got = render(engine, data)
    assert rel_err(got, g["photons"]) < REL                # vs the live reference
    # weights: exp() differs by an ulp between CUDA and glibc, so compare with the device's own
    dev_w = device_weights(engine, data, 0.033)
    assert abs(dev_w - weight).max() / weight.max() < 1e-15
    want = c_oracle.render_sat(c_oracle.geometry(params), data[:, 0], data[:, 1], data[:, 2], dev_w, sats, inv, slot)
    assert rel_err(got, want) < FIXED
    assert ((got > 0) == (want > 0)).all()                   # identical pixel footprints
    perm = numpy.random.RandomState(1).permutation(len(data))
    assert numpy.array_equal(render(engine, data[perm]), got)   # bit for bit, any spot order


def engine_keys(engine, configs, data):
    from scopyon_b200._epifm import depth_keys_of
    return depth_keys_of(data[:, 0] - configs.detector_focal_point[0], configs.depth_cutoff, engine.geom.n_depth_keys)


def device_weights(engine, data, unit_time):
    import ctypes
    from scopyon_b200 import _native
    soa = torch.from_numpy(numpy.ascontiguousarray(data[:, [0, 1, 2, 4]].T)).to(engine.device)
    w = torch.empty(len(data), dtype=torch.float64, device=engine.device)
    engine._call("scb_emit_bleach", 0, len(data), _native.ptr(soa[0]), _native.ptr(soa[1]), _native.ptr(soa[2]),
                 _native.ptr(soa[3]), None, None, float(unit_time), float(engine.configs.detector_focal_point[0]),
                 ctypes.byref(engine.phys), None, _native.ptr(w), None, engine._stream())
    return w.cpu().numpy()


def test_gaussian_case_matches_reference():
    g = golden("gaussian_case.npz")
    config, _, params, engine = gpu_engine("""
default:
    fluorophore: {type: Gaussian, radial_width: {value: 100.0e-9, units: m}, wave_length: {value: 600.0e-9, units: m}}
    detector: {image_size: [64, 64], exposure_time: 0.033}
""")
    got = render(engine, format_inputs(config, g["inputs"])[0][1])
    assert rel_err(got, g["photons"]) < REL


def test_large_random_scene_bit_exact_and_properties():
    """20 000 spots on 1024 x 1000 (ragged strips), a 3 000-spot cluster inside a few strips
    (a long work list for one warp), off-screen and zero-weight spots."""
    config, configs, params, engine = gpu_engine("""
default:
    detector: {type: CMOS, image_size: [1024, 1000], pixel_length: {value: 6.5e-6, units: m}, exposure_time: 0.033}
    magnification: 100
""")
    rng = numpy.random.RandomState(5)
    pl = configs.pixel_length
    n = 20000
    data = numpy.zeros((n, 5))
    data[:, 1] = rng.uniform(-540 * pl, 540 * pl, n)
    data[:, 2] = rng.uniform(-520 * pl, 520 * pl, n)
    data[:3000, 1] = rng.normal(100 * pl, 2 * pl, 3000)
    data[:3000, 2] = rng.normal(-200 * pl, 2 * pl, 3000)
    # spots on exact pixel centres / corners / whole nanometres: every edge expression lands on an
    # integer, where rounding decides the sample index (the unevenly spaced footprints of the SAT path)
    data[3000:3200, 1] = (rng.randint(-400, 400, 200) + 0.5 * rng.randint(0, 2, 200)) * pl
    data[3000:3200, 2] = (rng.randint(-400, 400, 200) + 0.5 * rng.randint(0, 2, 200)) * pl
    data[3200:3400, 1] = rng.randint(-20000, 20000, 200) * 1e-9
    data[3200:3400, 2] = rng.randint(-20000, 20000, 200) * 1e-9
    data[:, 3] = numpy.arange(n)
    data[:, 4] = (rng.uniform(size=n) > 0.1)                 # 10 % dark molecules
    sats, inv, slot = oracle_tables(params, engine, [0])
    got = render(engine, data)
    # 65 nm pixels: evenly spaced footprints come from the box table by TMA; without the box
    # table every footprint is summed from SAT corners -- the two paths agree bit for bit
    assert engine.box is not None
    box, engine.tables.box = engine.tables.box, None
    assert numpy.array_equal(render(engine, data), got)
    engine.tables.box = box
    dev_w = device_weights(engine, data, 0.033)
    assert (dev_w[data[:, 4] == 0] == 0).all()
    want = c_oracle.render_sat(c_oracle.geometry(params), data[:, 0], data[:, 1], data[:, 2], dev_w, sats, inv, slot)
    assert rel_err(got, want) < FIXED
    assert ((got > 0) == (want > 0)).all()
    perm = rng.permutation(n)
    assert numpy.array_equal(render(engine, data[perm]), got)   # the cluster too: no order dependence
    # linearity: rendering two halves separately and adding equals rendering all
    a = render(engine, data[: n // 2])
    b = render(engine, data[n // 2:])
    assert rel_err(a + b, got) < FIXED
    # mass: interior spots deposit weight * table integral
    interior = (abs(data[:, 1]) < 490 * pl) & (abs(data[:, 2]) < 480 * pl)
    only = render(engine, data[interior])
    assert abs(only.sum() - dev_w[interior].sum() * 0.9788254597277128) / only.sum() < 1e-11
    # run-to-run reproducibility (no atomics on the image)
    assert numpy.array_equal(render(engine, data[3000:]), render(engine, data[3000:]))
    # fp32 frames use an fp32 box table and 32-bit fixed-point accumulators whose LSB is set per strip
    # from the length of its list; this scene is the adverse case (a strip holding a 3 000-spot
    # cluster and faint background): 1e-6 of the image maximum (north_star asks 1e-5 for fp32).
    # Again both paths agree bit for bit
    _, _, _, engine32 = gpu_engine("""
default:
    detector: {type: CMOS, image_size: [1024, 1000], pixel_length: {value: 6.5e-6, units: m}, exposure_time: 0.033}
    magnification: 100
""", precision="f32")
    got32 = render(engine32, data, dtype=torch.float32)
    assert engine32.box.dtype == torch.float32
    assert rel_err(got32.astype(numpy.float64), got) < 3e-6
    import os
    os.environ["SCB_RENDER_FORCE_GATHER"] = "1"          # same tables and accumulators, SAT corners for every footprint
    try:
        assert numpy.array_equal(render(engine32, data, dtype=torch.float32), got32)
    finally:
        del os.environ["SCB_RENDER_FORCE_GATHER"]
    box, engine32.tables.box = engine32.tables.box, None   # no box table at all: exact 64-bit accumulation
    assert rel_err(render(engine32, data, dtype=torch.float32).astype(numpy.float64), got) < 3e-7
    engine32.tables.box = box


def test_empty_and_all_dark_inputs():
    config, _, params, engine = gpu_engine("default: {detector: {image_size: [40, 24]}}")
    assert (render(engine, numpy.zeros((0, 5))) == 0).all()
    data = numpy.zeros((5, 5))
    assert (render(engine, data) == 0).all()                  # p_state = 0 everywhere


def test_motion_blur_sums_snapshots():
    config, configs, params, engine = gpu_engine("default: {detector: {image_size: [64, 48], exposure_time: 0.03}}")
    rng = numpy.random.RandomState(2)
    pl = configs.pixel_length
    snaps = []
    for k in range(3):
        d = numpy.zeros((6, 5))
        d[:, 1:3] = rng.uniform(-20 * pl, 20 * pl, (6, 2))
        d[:, 3] = numpy.arange(6)
        d[:, 4] = 1
        snaps.append((0.01, d))
    out = torch.empty((64, 48), dtype=torch.float64, device=engine.device)
    img, true_data = engine.render_expected(snaps, out=out, want_true_data=True, exposure_time=0.03)
    want, want_true = orc.expected_frame([(0.01 * k, d) for k, (_, d) in enumerate(snaps)], params,
                                         exposure_time=0.03)
    assert rel_err(img.cpu().numpy(), want) < REL
    assert set(true_data) == set(want_true)
    for m in want_true:
        assert numpy.allclose(true_data[m], want_true[m], rtol=1e-13, atol=0)


@pytest.mark.parametrize("pixel_nm, magnification", [
    (44.0, 100),      # 48 slots per phase: wider than the TMA ring -> every footprint gathers SAT corners
    (65.0, 100),      # 32 slots: the box-table / TMA path at its compile-time width
    (100.0, 100),     # 24 slots: box-table path at run-time width, in fp32 too (slots are a multiple of four)
    (130.0, 100),     # 20 slots
    (160.0, 100),     # 16 slots, 14-pixel footprints: mostly idle lanes
    (250.0, 40),      # 10 slots, 9-pixel footprints
    (66.39, 241),     # default magnification: no whole number of samples per pixel, SAT path only
])
def test_pixel_pitches_against_c_oracle(pixel_nm, magnification):
    """Every table layout / kernel variant the pixel pitch selects gives the C oracle's image."""
    yaml = """
default:
    detector: {type: CMOS, image_size: [150, 210], pixel_length: {value: %.9e, units: m}, exposure_time: 0.033}
    magnification: %d
""" % (pixel_nm * 1e-9 * magnification, magnification)
    config, configs, params, engine = gpu_engine(yaml)
    pl = configs.pixel_length
    assert abs(pl / 1e-9 - pixel_nm) < 1e-6
    rng = numpy.random.RandomState(int(pixel_nm))
    n = 400
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 3, n).astype(int) * 150e-9             # three depth keys
    data[:, 1] = rng.uniform(-80 * pl, 80 * pl, n)
    data[:, 2] = rng.uniform(-110 * pl, 110 * pl, n)
    data[:40, 1:3] = numpy.round(data[:40, 1:3] / pl) * pl              # exact pixel centres
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    keys = numpy.unique(engine_keys(engine, configs, data))
    sats, inv, slot = oracle_tables(params, engine, keys)
    got = render(engine, data)
    dev_w = device_weights(engine, data, 0.033)
    want = c_oracle.render_sat(c_oracle.geometry(params), data[:, 0], data[:, 1], data[:, 2], dev_w, sats, inv, slot)
    assert rel_err(got, want) < FIXED
    assert ((got > 0) == (want > 0)).all()
    if engine.box is not None:       # and the SAT-corner path alone gives the same bits
        box, engine.tables.box = engine.tables.box, None
        assert numpy.array_equal(render(engine, data), got)
        engine.tables.box = box
    _, _, _, engine32 = gpu_engine(yaml, precision="f32")
    engine32.ensure_tables(keys)
    got32 = render(engine32, data, dtype=torch.float32)
    # 32-bit accumulators (box table present) or exact 64-bit ones (SAT only), fp32 output either way
    assert rel_err(got32.astype(numpy.float64), want) < (1e-6 if engine32.box is not None else 3e-7)


@pytest.mark.parametrize("pixel_nm, magnification, reg_kernel", [
    (65.0, 100, True),       # 32 slots per phase
    (44.0, 100, True),       # 48 slots: wider than the shared-memory kernel's ring, the register kernels do not mind
    (160.0, 100, True),      # 16 slots, 14-pixel footprints: the copy box overhangs the block on both axes (zero fill)
    (100.0, 100, True),      # 24 slots at run-time width
])
def test_register_kernels_equal_shared_memory_kernel(pixel_nm, magnification, reg_kernel):
    """fp32 mode: the register-accumulator kernels (SCB_RENDER_PATH=tensor: tensor-map TMA with zero-filled overhang;
    =ldg: plain loads), their SAT-corner gather and the default shared-memory kernel give the same bits, on ragged
    strips, clipped footprints, irregular footprints and a cluster."""
    import os
    yaml = """
default:
    detector: {type: CMOS, image_size: [150, 210], pixel_length: {value: %.9e, units: m}, exposure_time: 0.033}
    magnification: %d
""" % (pixel_nm * 1e-9 * magnification, magnification)
    config, configs, params, engine = gpu_engine(yaml, precision="f32")
    pl = configs.pixel_length
    rng = numpy.random.RandomState(int(pixel_nm) + 1)
    n = 1500
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 3, n).astype(int) * 150e-9
    data[:, 1] = rng.uniform(-90 * pl, 90 * pl, n)                      # some footprints hang over the image border
    data[:, 2] = rng.uniform(-120 * pl, 120 * pl, n)
    data[:300, 1] = rng.normal(10 * pl, 1.5 * pl, 300)                  # a cluster: one long strip list
    data[:300, 2] = rng.normal(-30 * pl, 1.5 * pl, 300)
    data[300:360, 1:3] = numpy.round(data[300:360, 1:3] / pl) * pl      # exact pixel centres: irregular footprints
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    engine.ensure_tables(numpy.unique(engine_keys(engine, configs, data)))
    assert engine.box is not None and engine.box.dtype == torch.float32
    from scopyon_b200 import _native
    assert (_native.load().scb_psf_sat_slots(engine.geom.n_radial, engine.geom.sat_modulus) % 4 == 0) == reg_kernel
    images = {}
    for name, env in [("default", {}), ("tensor", {"SCB_RENDER_PATH": "tensor"}), ("ldg", {"SCB_RENDER_PATH": "ldg"}),
                      ("tensor gather", {"SCB_RENDER_PATH": "tensor", "SCB_RENDER_FORCE_GATHER": "2"}),
                      ("ldg gather", {"SCB_RENDER_PATH": "ldg", "SCB_RENDER_FORCE_GATHER": "2"}),
                      ("shared gather", {"SCB_RENDER_FORCE_GATHER": "1"})]:
        os.environ.update(env)
        try:
            images[name] = render(engine, data, dtype=torch.float32)
        finally:
            for k in env:
                del os.environ[k]
    assert images["default"].max() > 0
    for name, image in images.items():
        assert numpy.array_equal(image, images["default"]), name
    exact = render(gpu_engine(yaml, precision="f64")[3], data)
    assert rel_err(images["default"].astype(numpy.float64), exact) < 3e-6
    # the CTA-tile kernel (32 x 128 tiles, one shared bulk copy per spot and tile) sets its accumulator LSB per tile:
    # equal to the strip kernels to that LSB, and as close to the exact image as they are
    for env in ({"SCB_RENDER_PATH": "tile"}, {"SCB_RENDER_PATH": "tile", "SCB_RENDER_FORCE_GATHER": "2"}):
        os.environ.update(env)
        try:
            tiled = render(engine, data, dtype=torch.float32)
        finally:
            for k in env:
                del os.environ[k]
        assert rel_err(tiled.astype(numpy.float64), images["default"].astype(numpy.float64)) < 2e-6
        assert rel_err(tiled.astype(numpy.float64), exact) < 3e-6
        assert ((tiled > 0) == (images["default"] > 0)).all()


@pytest.mark.parametrize("size, pixel_nm, n, why", [
    ((150, 210), 65.0, 1500, "a few dozen strips: the shared-memory census"),
    ((64, 8192), 65.0, 4000, "512 strips in one row of the box, scattered spots: shared memory, every counter touched"),
    ((4104, 1100), 65.0, 3000, "513 x 9 strips > 4096 counters: the global-atomic census takes over"),
    ((300, 300), 20.0, 600, "101-row footprints over more than 16 strips each: the global-atomic census"),
], ids=["small", "wide", "tall", "fine-pitch"])
def test_census_paths_give_the_same_image(size, pixel_nm, n, why):
    """spot_prepare counts a CTA's (spot, strip) overlaps in shared memory when the strips it touches fit 4096
    counters and its footprints 16 strips each, with global atomics otherwise (SCB_PREPARE_CENSUS=global forces
    that); strip_fill requests its list positions up front (SCB_FILL_VARIANT=old: strip by strip).  The lists
    come out in different orders; the image is the same, bit for bit (fixed-point accumulation)."""
    import os
    yaml = """
default:
    detector: {type: CMOS, image_size: [%d, %d], pixel_length: {value: %.9e, units: m}, exposure_time: 0.033}
    magnification: 100
""" % (size[0], size[1], pixel_nm * 1e-7)
    config, configs, params, engine = gpu_engine(yaml, precision="f32")
    pl = configs.pixel_length
    rng = numpy.random.RandomState(n)
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 3, n).astype(int) * 150e-9
    data[:, 1] = rng.uniform(-0.55 * size[0] * pl, 0.55 * size[0] * pl, n)      # some footprints hang over the border
    data[:, 2] = rng.uniform(-0.55 * size[1] * pl, 0.55 * size[1] * pl, n)
    data[: n // 5, 1] = rng.normal(0.1 * size[0] * pl, 1.5 * pl, n // 5)        # a cluster: one long strip list
    data[: n // 5, 2] = rng.normal(-0.2 * size[1] * pl, 1.5 * pl, n // 5)
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    engine.ensure_tables(numpy.unique(engine_keys(engine, configs, data)))
    images = {}
    for name, env in [("default", {}), ("global census", {"SCB_PREPARE_CENSUS": "global"}),
                      ("fill strip by strip", {"SCB_FILL_VARIANT": "old"}),
                      ("both", {"SCB_PREPARE_CENSUS": "global", "SCB_FILL_VARIANT": "old"})]:
        os.environ.update(env)
        try:
            images[name] = render(engine, data, dtype=torch.float32)
        finally:
            for k in env:
                del os.environ[k]
    assert images["default"].max() > 0, why
    for name, image in images.items():
        assert numpy.array_equal(image, images["default"]), (name, why)
    # and the list lengths add up: the image carries every spot's photons that fall inside the frame
    exact = render(gpu_engine(yaml, precision="f64")[3], data)
    assert rel_err(images["default"].astype(numpy.float64), exact) < 3e-6


def test_copy_engine_and_strip_shape_variants():
    """The measured variants of the shared-memory kernel -- per-lane cp.async instead of the TMA bulk copy
    (SCB_RENDER_COPY=ldgsts), units split between the two engines (=mix), 16 x 64 strips (SCB_RENDER_ROWS=16) --
    share its ring layout (a stage's mbarrier sits behind its rows).  Same strips: the same bits; other strips:
    another accumulator LSB per strip, equal to that LSB."""
    import os
    yaml = """
default:
    detector: {type: CMOS, image_size: [150, 210], pixel_length: {value: 6.5e-6, units: m}, exposure_time: 0.033}
    magnification: 100
"""
    config, configs, params, engine = gpu_engine(yaml, precision="f32")
    pl = configs.pixel_length
    rng = numpy.random.RandomState(11)
    n = 1500
    data = numpy.zeros((n, 5))
    data[:, 0] = rng.uniform(0, 3, n).astype(int) * 150e-9
    data[:, 1] = rng.uniform(-90 * pl, 90 * pl, n)
    data[:, 2] = rng.uniform(-120 * pl, 120 * pl, n)
    data[:300, 1] = rng.normal(10 * pl, 1.5 * pl, 300)
    data[:300, 2] = rng.normal(-30 * pl, 1.5 * pl, 300)
    data[300:360, 1:3] = numpy.round(data[300:360, 1:3] / pl) * pl      # exact pixel centres: irregular footprints
    data[:, 3] = numpy.arange(n)
    data[:, 4] = 1
    engine.ensure_tables(numpy.unique(engine_keys(engine, configs, data)))

    def image(env):
        os.environ.update(env)
        try:
            return render(engine, data, dtype=torch.float32)
        finally:
            for k in env:
                del os.environ[k]

    default = image({})
    assert default.max() > 0
    for env in ({"SCB_RENDER_COPY": "ldgsts"}, {"SCB_RENDER_COPY": "mix"}):
        assert numpy.array_equal(image(env), default), env
    for env in ({"SCB_RENDER_ROWS": "16"}, {"SCB_RENDER_ROWS": "16", "SCB_RENDER_COPY": "ldgsts"}):
        tall = image(env)
        assert rel_err(tall.astype(numpy.float64), default.astype(numpy.float64)) < 2e-6, env
        assert ((tall > 0) == (default > 0)).all(), env
