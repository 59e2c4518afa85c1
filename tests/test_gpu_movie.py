"""Device-resident movies (scopyon_b200.movie.DeviceMovie): frame-block sharding must
reproduce the sequential movie exactly, and frames must agree with the oracle."""
import numpy
import pytest
import torch

import epifm_oracle as orc
import scopyon_b200
from conftest import make_configs
from scopyon_b200.movie import DeviceMovie, frame_block

pytestmark = pytest.mark.gpu

YAML = """
default:
    magnification: 100
    fluorophore: {depth_cutoff: {value: 200.0e-9, units: m}}     # 203 depth tables instead of 1003: quick to build
    light_source: {angle: {value: %s, units: radian}}
    detector: {type: CMOS, image_size: [96, 80], pixel_length: {value: 6.5e-6, units: m}, QE: 0.73, exposure_time: 0.033}
    analog_to_digital_converter: {bit: 16, offset: 100, fullwell: 30000, type: column, count: 2.0}
    effects: {photo_bleaching: {switch: %s, half_life: {value: 0.08, units: s}}}
"""


def make_movie(angle, bleaching, n=400, seed=11, precision="f32", pixel_length=None):
    config = scopyon_b200.DefaultConfiguration()
    config.update(YAML % (angle, bleaching))
    if pixel_length is not None:
        config.default.detector.pixel_length = pixel_length
    pl = 6.5e-8
    lower, upper = [-48 * pl, -40 * pl, 0.0], [48 * pl, 40 * pl, 1.3e-6]
    return config, DeviceMovie(config, n, lower, upper, 2e-13, seed, precision=precision)


@pytest.mark.parametrize("angle", ["0.0", "1.2566370614359172"])      # epi, TIRF (depth-dependent emission)
def test_frame_blocks_reproduce_the_sequential_movie(angle):
    _, whole = make_movie(angle, "true")
    frames = torch.empty((7, 96, 80), dtype=torch.float32, device=whole.engine.device)
    positions = []
    for k in range(7):
        positions.append(whole.positions())
        whole.render_next(frames[k])
    frames = frames.cpu().numpy()
    budgets = whole.budget.cpu().numpy()
    assert (budgets == 0).sum() > 20 and (budgets > 0).sum() > 20     # bleaching is at work

    world = 3
    for rank in range(world):
        first, last = frame_block(7, rank, world)
        _, part = make_movie(angle, "true")
        part.reset(first_frame=first)                                   # replay, no rendering
        assert numpy.array_equal(part.positions(), positions[first])
        block = torch.empty((last - first, 96, 80), dtype=torch.float32, device=part.engine.device)
        part.render_block(block)
        assert numpy.array_equal(block.cpu().numpy(), frames[first:last])   # bit for bit
    other = make_movie(angle, "true", seed=12)[1]
    one = torch.empty((1, 96, 80), dtype=torch.float32, device=other.engine.device)
    other.render_block(one)
    assert not numpy.array_equal(one.cpu().numpy()[0], frames[0])


def test_movie_frame_matches_oracle_expectation():
    config, movie = make_movie("0.0", "false", precision="f64")
    _, _, params = make_configs(YAML % ("0.0", "false"))
    for _ in range(2):
        data = movie.positions()
        adc = torch.empty((96, 80), dtype=torch.float64, device=movie.engine.device)
        expectation = torch.empty_like(adc)
        movie.render_next(adc, expectation_out=expectation)
        photons, _ = orc.expected_frame([(0.0, data)], params, exposure_time=0.033)
        want = orc.detector_expectation(photons, params)
        got = expectation.cpu().numpy()
        assert abs(got - want).max() / want.max() < 1e-9
    # Brownian step between the two frames: displacement variance 2 D dt per axis
    step = movie.positions()[:, :3] - data[:, :3]
    assert abs(step.std(axis=0) / numpy.sqrt(2 * 2e-13 * 0.033) - 1).max() < 0.15


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_gather_frames_nccl_two_gpus(tmp_path):
    import os
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = numpy.load(os.path.join(str(tmp_path), "stack0.npy"))
    b = numpy.load(os.path.join(str(tmp_path), "stack1.npy"))
    assert a.shape == (5, 96, 80) and numpy.array_equal(a, b)
    _, whole = make_movie("0.0", "true")
    frames = torch.empty((5, 96, 80), dtype=torch.float32, device=whole.engine.device)
    whole.render_block(frames)
    assert numpy.array_equal(frames.cpu().numpy(), a)


def _nccl_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    from scopyon_b200.movie import gather_frames
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    first, last = frame_block(5, rank, world)
    _, movie = make_movie("0.0", "true")
    movie.reset(first_frame=first)
    block = torch.empty((last - first, 96, 80), dtype=torch.float32, device=movie.engine.device)
    movie.render_block(block)
    stack = gather_frames(block, 5)
    numpy.save(os.path.join(out_dir, "stack{}.npy".format(rank)), stack.cpu().numpy())
    dist.destroy_process_group()


def test_frames_to_8bit_matches_image_as_8bit():
    """Device-side 8-bit scaling of a frame stack = Image.as_8bit per frame with the common
    limits Video.save uses (image.py:98-123, 261-264), bit for bit."""
    import scopyon_b200
    from scopyon_b200.movie import frames_to_8bit
    rng = numpy.random.RandomState(11)
    stack = (rng.gamma(2.0, 300.0, (3, 37, 53)) + 100).astype(numpy.float32)
    stack[1, 5, 7] = 65535.0
    dev = torch.from_numpy(stack).cuda()
    for kwargs in ({}, {"cmin": 150.0, "cmax": 2000.0}, {"low": 10, "high": 200}, {"cmin": 0.0}):
        got = frames_to_8bit(dev, **kwargs).cpu().numpy()
        limits = dict(kwargs)
        limits.setdefault("cmin", float(stack.min()))
        limits.setdefault("cmax", float(stack.max()))
        want = numpy.stack([scopyon_b200.Image(f.astype(numpy.float64)).as_8bit(**limits).as_array() for f in stack])
        assert got.dtype == numpy.uint8 and numpy.array_equal(got, want)
    flat = torch.full((2, 8, 8), 7.0, device="cuda")
    assert (frames_to_8bit(flat, low=3).cpu().numpy() == 3).all()          # cmax == cmin -> low
    f64 = torch.from_numpy(stack.astype(numpy.float64)).cuda()
    assert numpy.array_equal(frames_to_8bit(f64).cpu().numpy(), frames_to_8bit(dev).cpu().numpy())


@pytest.mark.parametrize("frames_per_launch", [16, 3])
def test_stream_frames_exports_the_movie_block_by_block(frames_per_launch, tmp_path):
    """The data plane of a sharded movie: stream_frames hands every finished block to the host
    (float32 exact; uint16 = rint/clip; uint8 = Image.as_8bit with fixed limits) and the frames are
    the ones render_block produces, bit for bit; save_movie writes the same stack as a .npy file."""
    import scopyon_b200
    from scopyon_b200.movie import save_movie
    n_frames = 23                                         # ragged last block
    _, whole = make_movie("0.0", "true")
    frames = torch.empty((n_frames, 96, 80), dtype=torch.float32, device=whole.engine.device)
    whole.render_block(frames)
    frames = frames.cpu().numpy()

    def collect(fmt, **kwargs):
        _, movie = make_movie("0.0", "true")
        movie.export_block_frames = frames_per_launch
        got, firsts = [], []

        def sink(first, block):
            firsts.append(first)
            got.append(block.copy())                      # the view is only valid during the call
        assert movie.stream_frames(n_frames, fmt=fmt, sink=sink, ring=2, **kwargs) == n_frames
        assert firsts == list(range(0, n_frames, frames_per_launch)) and movie.frame == n_frames
        return numpy.concatenate(got)

    f32 = collect("f32")
    assert f32.dtype == numpy.float32 and numpy.array_equal(f32, frames)
    u16 = collect("u16")
    assert u16.dtype == numpy.uint16
    assert numpy.array_equal(u16, numpy.rint(numpy.clip(frames.astype(numpy.float64), 0, 65535)).astype(numpy.uint16))
    u8 = collect("u8", limits=(90.0, 140.0))
    want = numpy.stack([scopyon_b200.Image(f.astype(numpy.float64)).as_8bit(cmin=90.0, cmax=140.0).as_array()
                        for f in frames])
    assert u8.dtype == numpy.uint8 and numpy.array_equal(u8, want)
    # a second call continues the movie where the first stopped
    _, movie = make_movie("0.0", "true")
    movie.export_block_frames = frames_per_launch
    parts = []
    movie.stream_frames(10, sink=lambda first, block: parts.append(block.copy()))
    movie.stream_frames(n_frames - 10, sink=lambda first, block: parts.append(block.copy()))
    assert numpy.array_equal(numpy.concatenate(parts), frames)
    # straight to disk
    _, movie = make_movie("0.0", "true")
    assert save_movie(movie, tmp_path / "movie.npy", n_frames) == n_frames
    assert numpy.array_equal(numpy.load(tmp_path / "movie.npy"), frames)
    _, movie = make_movie("0.0", "true")
    save_movie(movie, tmp_path / "movie16.npy", n_frames, fmt="u16")
    assert numpy.array_equal(numpy.load(tmp_path / "movie16.npy"), u16)


def test_frames_to_u16():
    rng = numpy.random.RandomState(5)
    stack = rng.uniform(-10, 70000, (2, 33, 41)).astype(numpy.float32)
    stack[0, 0, :6] = [0.5, 1.5, 2.5, -0.0, 65535.4, numpy.nan]        # ties to even, clip, NaN -> 0
    lib = scopyon_b200._native.load()
    for dtype in (torch.float32, torch.float64):
        dev = torch.from_numpy(stack).cuda().to(dtype)
        out = torch.empty(stack.shape, dtype=torch.uint16, device="cuda")
        assert lib.scb_frames_to_u16(dev.data_ptr(), dev.numel(), 0 if dtype == torch.float32 else 1, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        want = numpy.rint(numpy.clip(numpy.nan_to_num(stack.astype(numpy.float64), nan=0.0), 0, 65535)).astype(numpy.uint16)
        assert numpy.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("pixel_length", [None, 6.639e-6], ids=["box-tables", "sat-corners"])
@pytest.mark.parametrize("tight", [False, True], ids=["plan", "tight-plan"])
def test_planned_blocks_equal_unplanned_blocks(tight, pixel_length):
    """Consecutive blocks of a movie bin their spots in one pass over the list plan the previous block left behind
    (scb_render_expected_frames_planned); a plan that is too small (the test hook halves every list's room) sends
    units to the overflow list.  Either way the frames are those of the count / scan / fill pipeline, bit for bit."""
    import os
    n_frames, nf = 40, 8

    def movie_frames(plan):
        _, movie = make_movie("0.0", "true", n=3000, pixel_length=pixel_length)     # 66.39 nm: every footprint walks its edges
        movie.plan_blocks = plan
        movie.frames_per_launch = nf
        frames = torch.empty((n_frames, 96, 80), dtype=torch.float32, device=movie.engine.device)
        movie.render_block(frames)
        torch.cuda.synchronize()
        assert int(movie.engine.errors.item()) == 0
        return frames.cpu().numpy()

    want = movie_frames(False)
    if tight:
        os.environ["SCB_PLAN_TIGHT"] = "1"
    try:
        got = movie_frames(True)
    finally:
        os.environ.pop("SCB_PLAN_TIGHT", None)
    assert want.max() > 0 and numpy.array_equal(got, want)


def test_a_plan_that_falls_short_by_too_much_is_reported():
    """More units beyond their strips' room than the overflow list holds (65 536): the device counts them and the
    movie raises -- at the next block, or when asked (check_errors) -- instead of handing out frames with spots missing."""
    import os
    _, movie = make_movie("0.0", "false", n=12000)
    movie.frames_per_launch = 8
    frames = torch.empty((16, 96, 80), dtype=torch.float32, device=movie.engine.device)
    os.environ["SCB_PLAN_TIGHT"] = "1"
    try:
        movie.render_block(frames[:8])          # first block: count / scan / fill, leaves a plan of half the room
        movie.check_errors()
        movie.render_block(frames[8:])          # ~ 3e5 units, half of them beyond their strip's room
        with pytest.raises(RuntimeError, match="planned block"):
            movie.check_errors()
    finally:
        os.environ.pop("SCB_PLAN_TIGHT", None)
        movie.engine.errors.zero_()
