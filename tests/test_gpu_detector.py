"""Detector + ADC pass on the GPU (C ABI: scb_detector_adc, scb_adc_offsets).

ADC arithmetic is checked bit for bit with injected draws; the samplers are checked
statistically against the reference's distributions (moments, KS / chi-square at the
stated sample sizes)."""
import numpy
import pytest
import scipy.stats
import torch

import epifm_oracle as orc
from conftest import cmos_table, golden, gpu_engine

pytestmark = pytest.mark.gpu

CCD = "default: {detector: {type: CCD, image_size: [256, 192], readout_noise: 3.0}}"


def run_detector(engine, photons, frame=0, seed=1234, taps=True, **inject):
    dev = photons.device
    adc = torch.empty_like(photons)
    expectation = torch.empty_like(photons)
    sig = torch.empty_like(photons) if taps else None
    noi = torch.empty_like(photons) if taps else None
    engine.detect(photons, frame, seed, adc=adc, expectation=expectation, out_signal=sig, out_noise=noi, **inject)
    torch.cuda.synchronize()
    f = lambda t: None if t is None else t.cpu().numpy()
    return f(adc), f(expectation), f(sig), f(noi)


@pytest.mark.parametrize("fpn", ["none", "column", "pixel"])
def test_adc_bit_exact_with_injected_draws(fpn):
    yaml = """
default:
    detector: {type: CCD, image_size: [60, 44], readout_noise: 3.0}
    analog_to_digital_converter: {type: %s, count: 3.0, offset: 100, fullwell: 30000}
""" % fpn
    _, configs, params, engine = gpu_engine(yaml, precision="f64")
    rng = numpy.random.RandomState(3)
    shape = (60, 44)
    photons = rng.uniform(0, 50, shape)
    signal = rng.poisson(40.0, shape).astype(float)
    signal[0, :5] = [0, 1e6, 29999.5, 30000.5, 123.4]
    noise = rng.normal(0, 80, shape)                       # large: drives some pixels negative / above full well
    to = lambda a: torch.from_numpy(a).to(engine.device)
    adc, expectation, _, _ = run_detector(engine, to(photons), in_signal=to(signal), in_noise=to(noise))
    if fpn == "none":
        offset, gain = orc.adc_params(params)
    else:
        dev_offset = engine.offset.cpu().numpy()
        assert (dev_offset == numpy.rint(dev_offset)).all()
        normals = dev_offset            # already rint-ed; rint is idempotent
        offset, gain = orc.adc_params(params, normals)
    want = orc.adc_counts(signal + noise, params["adc_fullwell"], gain, offset, params["adc_bit"])
    assert numpy.array_equal(adc, want)                                    # bit for bit
    assert numpy.array_equal(expectation, orc.detector_expectation(photons, params))
    assert adc.min() >= 0 and adc.max() <= 2 ** 16 - 1


def test_adc_known_answers(known_answers):
    _, configs, params, engine = gpu_engine("default: {detector: {type: CCD, image_size: [1, 8]}}", precision="f64")
    pe = numpy.array(known_answers["adc"]["pe"] + [0.0])
    to = lambda a: torch.from_numpy(a.reshape(1, 8)).to(engine.device)
    adc, _, _, _ = run_detector(engine, to(numpy.zeros(8)), in_signal=to(pe), in_noise=to(numpy.zeros(8)))
    assert adc.ravel()[:7].tolist() == known_answers["adc"]["counts"]


def test_fpn_offsets_distribution():
    yaml = """
default:
    detector: {type: CMOS, image_size: [512, 512]}
    analog_to_digital_converter: {type: pixel, count: 2.0, offset: 100}
"""
    _, _, _, engine = gpu_engine(yaml, precision="f32")
    off = engine.offset.cpu().numpy().astype(float)
    assert (off == numpy.rint(off)).all()
    assert abs(off.mean() - 100) < 0.02 and abs(off.std() - numpy.sqrt(4 + 1 / 12.0)) < 0.02
    _, _, _, again = gpu_engine(yaml, precision="f32")
    assert numpy.array_equal(off, again.offset.cpu().numpy())               # same rng seed -> same map


@pytest.mark.parametrize("lam", [0.0092, 0.7, 5.0, 11.9, 12.1, 30.0, 80.0, 600.0, 2040.0, 2500.0])
def test_poisson_shot_noise(lam):
    """CCD/CMOS signal ~ Poisson(E) (_epifm.py:352,432): moments + KS on 49 152 pixels.  E < 12: inversion;
    12 <= E < 2048: first PTRS trial in the streaming kernel, continued in the second pass when undecided;
    beyond: PTRS in the second pass."""
    _, _, params, engine = gpu_engine(CCD, precision="f32")
    qe, bg = params["QE"], params["background_mean"]
    photons = torch.full((256, 192), lam / qe - bg, dtype=torch.float32, device=engine.device)
    _, expectation, sig, _ = run_detector(engine, photons)
    lam_eff = float(expectation.astype(float).mean())
    n = sig.size
    assert (sig == numpy.rint(sig)).all() and sig.min() >= 0
    assert abs(sig.mean() - lam_eff) < 5 * numpy.sqrt(lam_eff / n)
    assert abs(sig.var() / lam_eff - 1) < 5 * numpy.sqrt(2.0 / n) + 3 / (lam_eff * n) ** 0.5
    # discrete KS: compare empirical cdf to the Poisson cdf on the integers
    ks = numpy.arange(0, int(sig.max()) + 2)
    ecdf = numpy.searchsorted(numpy.sort(sig.ravel()), ks, side="right") / n
    assert abs(ecdf - scipy.stats.poisson.cdf(ks, lam_eff)).max() < 1.63 / numpy.sqrt(n)   # alpha = 0.01


def test_gaussian_readout_noise_and_determinism():
    _, _, params, engine = gpu_engine(CCD, precision="f32")
    photons = torch.zeros((256, 192), dtype=torch.float32, device=engine.device)
    adc, _, sig, noi = run_detector(engine, photons, frame=3)
    n = noi.size
    assert abs(noi.mean()) < 5 * 3.0 / numpy.sqrt(n) and abs(noi.std() / 3.0 - 1) < 5 / numpy.sqrt(2 * n)
    assert scipy.stats.kstest(noi.ravel() / 3.0, "norm").pvalue > 1e-3
    assert abs(numpy.corrcoef(noi[:, :-1].ravel(), noi[:, 1:].ravel())[0, 1]) < 5 / numpy.sqrt(n)
    again, _, _, _ = run_detector(engine, photons, frame=3)
    other, _, _, _ = run_detector(engine, photons, frame=4)
    assert numpy.array_equal(adc, again) and not numpy.array_equal(adc, other)
    # counter-based: a crop of the frame draws the same numbers as the full frame's head
    head = torch.zeros((64, 192), dtype=torch.float32, device=engine.device)
    yaml = "default: {detector: {type: CCD, image_size: [64, 192], readout_noise: 3.0}}"
    _, _, _, small = gpu_engine(yaml, precision="f32")
    crop, _, _, _ = run_detector(small, head, frame=3)
    assert numpy.array_equal(crop, adc[:64])


def test_cmos_readout_noise_matches_table():
    """CMOS noise is a categorical draw from RNDist_F40 (_epifm.py:334-345): chi-square."""
    _, _, params, engine = gpu_engine("default: {detector: {type: CMOS, image_size: [1024, 1024]}}", precision="f32")
    photons = torch.zeros((1024, 1024), dtype=torch.float32, device=engine.device)
    _, _, sig, noi = run_detector(engine, photons)
    values, p = orc.cmos_readout_pmf(cmos_table())
    idx = numpy.rint((noi.ravel().astype(float) - values[0]) / 0.1).astype(int)
    assert idx.min() >= 0 and idx.max() < len(values)
    assert numpy.allclose(values[idx], noi.ravel(), atol=1e-5)
    counts = numpy.bincount(idx, minlength=len(values)).astype(float)
    n = noi.size
    keep = p * n >= 10
    chi2 = ((counts[keep] - p[keep] * n) ** 2 / (p[keep] * n)).sum()
    assert chi2 < scipy.stats.chi2.ppf(0.999, keep.sum() - 1)
    assert counts[~keep].sum() <= max(20.0, 5 * p[~keep].sum() * n)
    assert abs(noi.mean() - (values * p).sum()) < 5 * numpy.sqrt(((values ** 2 * p).sum()) / n)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_emccd_signal_matches_reference_pmf(case):
    """EMCCD.get_signal draws from the pmf of _epifm.py:365-391.  The golden file holds the
    reference's own cdf for E = 0.0092 (background level), 0.5, 5 and 50 at gain 300."""
    g = golden("emccd_pmf.npz")
    E = float(g["E{}".format(case)])
    _, _, params, engine = gpu_engine("default: {detector: {image_size: [512, 512]}}", precision="f32")
    qe, bg = params["QE"], params["background_mean"]
    photons = torch.full((512, 512), E / qe - bg, dtype=torch.float32, device=engine.device)
    _, expectation, sig, _ = run_detector(engine, photons)
    sig = sig.ravel().astype(float)
    n = sig.size
    lo, hi = g["support{}".format(case)]
    assert sig.min() >= lo and sig.max() <= hi and (sig == numpy.rint(sig)).all()
    S, cdf = g["S{}".format(case)].astype(float), g["cdf{}".format(case)]
    ecdf = numpy.searchsorted(numpy.sort(sig), S, side="right") / n
    assert abs(ecdf - cdf).max() < 1.63 / numpy.sqrt(n) + 2e-3             # KS, alpha = 0.01 (+ grid thinning)
    mean, var = float(g["mean{}".format(case)]), float(g["var{}".format(case)])
    assert abs(sig.mean() - mean) < 5 * numpy.sqrt(var / n) + 1e-3 * mean
    assert abs(sig.var() / var - 1) < 0.05
    if lo == 0:
        p0 = float(g["p0_{}".format(case)])
        assert abs((sig == 0).mean() - p0) < 5 * numpy.sqrt(p0 * (1 - p0) / n) + 1e-4


def test_f32_and_f64_paths_agree_on_adc():
    yaml = "default: {detector: {type: CCD, image_size: [64, 64], readout_noise: 2.0}}"
    _, _, _, e32 = gpu_engine(yaml, precision="f32")
    _, _, _, e64 = gpu_engine(yaml, precision="f64")
    rng = numpy.random.RandomState(0)
    photons = rng.uniform(0, 30, (64, 64))
    a32, _, s32, n32 = run_detector(e32, torch.from_numpy(photons.astype(numpy.float32)).to(e32.device))
    a64, _, s64, n64 = run_detector(e64, torch.from_numpy(photons).to(e64.device))
    assert (s32 == s64).mean() > 0.999 and numpy.allclose(n32, n64, rtol=1e-6, atol=1e-6)
    assert numpy.allclose(a32, a64, rtol=2e-6, atol=1e-3)


@pytest.mark.parametrize("det", ["CMOS", "CCD", "EMCCD"])
@pytest.mark.parametrize("fpn", ["none", "column", "pixel"])
def test_streaming_kernel_equals_generic_kernel(det, fpn):
    """The production path (fp32, no taps: detector_fast_kernel + slow-pixel pass) and the
    generic kernel (taps requested) draw from the same Philox streams with the same
    arithmetic: identical ADC counts, pixel for pixel, dim and bright pixels alike."""
    yaml = """
default:
    detector: {type: %s, image_size: [192, 160], readout_noise: 2.5}
    analog_to_digital_converter: {type: %s, count: 3.0, offset: 100, fullwell: 30000}
""" % (det, fpn)
    _, _, _, engine = gpu_engine(yaml, precision="f32")
    rng = numpy.random.RandomState(11)
    photons = rng.exponential(1.5, (192, 160))
    photons[:40] = rng.exponential(9.0, (40, 160))          # around the small/large Poisson switch
    photons[40:48] = rng.uniform(50, 4000, (8, 160))        # bright rows: general samplers
    photons[48, :4] = [0.0, 1e-6, 60000.0, 1e7]             # dark, nearly dark, beyond full well
    photons = torch.from_numpy(photons.astype(numpy.float32)).to(engine.device)
    for frame in (0, 7):
        generic, _, sig, _ = run_detector(engine, photons, frame=frame, seed=99)
        fast = torch.empty_like(photons)
        engine.detect(photons, frame, 99, adc=fast)
        torch.cuda.synchronize()
        assert numpy.array_equal(fast.cpu().numpy(), generic)
    assert sig.max() > 100 and sig.min() == 0


def test_poisson_inversion_never_returns_the_loop_bound():
    """The inversion sampler (expectation < 12 e-) compares a 32-bit word with an fp32 running sum that
    stops moving a few hundred below 2^32; the largest words must map to the last count that could still
    be told apart, not to the search bound (a 160 e- hot pixel with probability ~1e-7 per pixel).
    The map word -> count is monotone and equals the exact quantile function away from its steps."""
    from scopyon_b200 import _native
    lib = _native.load()
    lambdas = numpy.array([1e-6, 0.0092, 0.05, 0.3, 0.7, 1.0, 1.9, 2.5, 3.3, 4.0, 5.7, 7.7, 9.9, 11.0, 11.999],
                          dtype=numpy.float32)
    words = numpy.array([0xffffffff, 0xfffffffe, 0xffffff80, 0xffffff00, 0xfffffe80, 0xfffffe00, 0xfffffd00,
                         0xfffff000, 0xffff0000, 0xff000000, 0x80000000, 0x1000, 0], dtype=numpy.uint32)
    lam, wrd = [a.ravel() for a in numpy.meshgrid(lambdas, words, indexing="ij")]
    # +- 1 ulp around every expectation as well (the saturation depends on the rounding of the sum)
    lam = numpy.concatenate([lam, numpy.nextafter(lam, numpy.float32(0)), numpy.nextafter(lam, numpy.float32(100))])
    wrd = numpy.concatenate([wrd, wrd, wrd])
    d_lam = torch.from_numpy(numpy.ascontiguousarray(lam, dtype=numpy.float32)).cuda()
    d_wrd = torch.from_numpy(wrd.astype(numpy.int64)).cuda().to(torch.uint32)
    out = torch.empty(len(lam), dtype=torch.float32, device="cuda")
    assert lib.scb_test_poisson_inversion(len(lam), d_lam.data_ptr(), d_wrd.data_ptr(), out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    counts = out.cpu().numpy()
    # a count beyond the 1 - 2^-33 quantile (+2) can only come from a stuck search
    ceiling = scipy.stats.poisson.isf(2.0 ** -33, lam.astype(numpy.float64)) + 2
    assert (counts <= ceiling).all(), (lam[counts > ceiling], wrd[counts > ceiling], counts[counts > ceiling])
    assert counts.max() < 60
    # the largest word gives the largest count of its expectation; word 0 gives 0 for every expectation
    by_lambda = counts.reshape(3, len(lambdas), len(words))
    assert (numpy.diff(by_lambda, axis=2) <= 0).all() and (by_lambda[:, :, -1] == 0).all()
    # random words: the exact quantile function except within rounding distance of a step
    rng = numpy.random.RandomState(1)
    n = 1 << 20
    lam = rng.choice(lambdas[1:], n).astype(numpy.float32)
    wrd = rng.randint(0, 2 ** 32, n, dtype=numpy.uint64).astype(numpy.uint32)
    d_lam = torch.from_numpy(lam).cuda()
    d_wrd = torch.from_numpy(wrd.astype(numpy.int64)).cuda().to(torch.uint32)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    assert lib.scb_test_poisson_inversion(n, d_lam.data_ptr(), d_wrd.data_ptr(), out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    u = (wrd.astype(numpy.float64) + 0.5) / 2.0 ** 32
    want = scipy.stats.poisson.ppf(u, lam.astype(numpy.float64))
    # the sampler's cumulative distribution is the exact one to ~1e-5 relative (a scale of 1 + 2.6e-6, see
    # detector.cu, the fp32 exponent and recurrence): every count must be the exact quantile of a word
    # within that distance, and only a small fraction of the words may land on a neighbouring count at all
    got = out.cpu().numpy()
    lam64 = lam.astype(numpy.float64)
    low = scipy.stats.poisson.ppf(u * (1 - 2e-5), lam64)
    high = scipy.stats.poisson.ppf(numpy.minimum(u * (1 + 2e-5), 1 - 2.0 ** -40), lam64)
    assert ((got >= low) & (got <= high)).all()
    assert (got != want).mean() < 1e-3


@pytest.mark.parametrize("fpn", ["none", "column", "pixel"])
def test_production_fp32_adc_against_oracle_on_its_own_draws(fpn):
    """The streaming kernel (fp32, reciprocal-multiply ADC on packed pairs) against the oracle's ADC
    formula (_epifm.py:1472-1484) evaluated in float64 on the SAME draws: signal and noise of every pixel
    are tapped from the generic kernel, which draws them from the same Philox words, and the counts the
    streaming kernel wrote must equal orc.adc_counts(signal + noise) to fp32 rounding."""
    yaml = """
default:
    detector: {type: CMOS, image_size: [256, 192], QE: 0.73}
    analog_to_digital_converter: {type: %s, count: 3.0, offset: 100, fullwell: 30000}
""" % fpn
    _, configs, params, engine = gpu_engine(yaml, precision="f32")
    rng = numpy.random.RandomState(8)
    photons = rng.exponential(3.0, (256, 192)).astype(numpy.float32)
    photons[:4] = rng.uniform(20000, 60000, (4, 192))          # up to and beyond the full well
    dev = torch.from_numpy(photons).to(engine.device)
    fast = torch.empty_like(dev)
    engine.detect(dev, 3, 2024, adc=fast)                       # production path: no taps
    _, _, signal, noise = run_detector(engine, dev, frame=3, seed=2024)
    torch.cuda.synchronize()
    if fpn == "none":
        offset, gain = orc.adc_params(params)
    else:
        offset, gain = orc.adc_params(params, engine.offset.cpu().numpy().astype(numpy.float64))
    want = orc.adc_counts(signal.astype(numpy.float64) + noise.astype(numpy.float64), params["adc_fullwell"], gain,
                          offset, params["adc_bit"])
    got = fast.cpu().numpy().astype(numpy.float64)
    assert abs(got - want).max() <= 4e-7 * want.max()           # a few fp32 ulps of the count
    assert got.max() == 2 ** 16 - 1 and got.min() >= 0 and (signal > 12).sum() > 500


def test_bright_frame_poisson_on_a_million_pixels():
    """A frame in which every pixel is bright (the first PTRS trial is made where the pixel is streamed and
    86-92 % of the pixels never reach the second pass): a tighter KS than test_poisson_shot_noise, on 1 048 576
    pixels per level, and the composite sampler's mean / variance / skewness."""
    yaml = "default:\n    detector: {type: CMOS, image_size: [1024, 1024], QE: 0.73}\n"
    _, _, params, engine = gpu_engine(yaml, precision="f32")
    qe, bg = params["QE"], params["background_mean"]
    for lam in (14.0, 57.3, 431.0):
        photons = torch.full((1024, 1024), lam / qe - bg, dtype=torch.float32, device=engine.device)
        _, expectation, sig, _ = run_detector(engine, photons, seed=5)
        lam_eff = float(expectation.astype(float).mean())
        n = sig.size
        assert (sig == numpy.rint(sig)).all() and sig.min() >= 0
        assert abs(sig.mean() - lam_eff) < 5 * numpy.sqrt(lam_eff / n)
        assert abs(sig.var() / lam_eff - 1) < 5 * numpy.sqrt(2.0 / n) + 3 / (lam_eff * n) ** 0.5
        skew = ((sig - sig.mean()) ** 3).mean() / sig.std() ** 3
        assert abs(skew - lam_eff ** -0.5) < 5 * numpy.sqrt(6.0 / n)
        ks = numpy.arange(0, int(sig.max()) + 2)
        ecdf = numpy.searchsorted(numpy.sort(sig.ravel()), ks, side="right") / n
        assert abs(ecdf - scipy.stats.poisson.cdf(ks, lam_eff)).max() < 1.63 / numpy.sqrt(n)   # alpha = 0.01
        # the production path gives the same counts as the generic kernel that returned `sig`
        fast = torch.empty_like(photons)
        generic, _, _, _ = run_detector(engine, photons, seed=5)
        engine.detect(photons, 0, 5, adc=fast)
        torch.cuda.synchronize()
        assert numpy.array_equal(fast.cpu().numpy(), generic)
