"""Blob and spot detection on the device, with the reference's interface
(``/root/reference/src/scopyon/analysis/spot_detection.py``).

``blob_detection`` is the reference's call into ``skimage.feature.blob_log`` (``:14-47``):
the Laplacian-of-Gaussian scale space and its 3x3x3 peaks are computed by
``scb_log_scale_space`` / ``scb_log_peaks``; the handful of surviving peaks are ordered and
pruned on the host exactly as scikit-image does it (``_prune_blobs``).  ``spot_detection``
(``:139-174``) fits every blob on the GPU (``scb_spot_fit``: one warp per blob).

The image may be a numpy array or a CUDA tensor (e.g. a frame that never left the device).
"""
import ctypes
import math
from logging import getLogger

import numpy

from .. import _native

_log = getLogger(__name__)

__all__ = ["blob_detection", "spot_detection"]

_SKIP_REASONS = {
    1: "spot_detection skip a blob due to the low signal.",
    2: "spot_detection skip a blob due to the failure in background.",
    3: "spot_detection skip a blob due to the failure in fitgaussian.",
    4: "spot_detection skip a blob due to invalid parameters fitted.",
}

MAX_FIT_ITERATIONS = 400        # scipy's least_squares allows 100 * n = 400 evaluations
SCALE_SPACE_BYTES = 1 << 30     # scratch bound: the cube is built this many bytes of scales at a time


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _native.NativeError("scopyon_b200.analysis needs a CUDA device (there is no CPU fallback)")
    return torch


def _as_float_image(data):
    """``skimage.util.img_as_float`` for the dtypes an image comes in: floats pass through,
    unsigned integers are divided by their type's maximum."""
    data = numpy.asarray(data)
    if data.ndim != 2:
        raise ValueError("a 2-D image is required; shape {} was given".format(data.shape))
    if data.dtype.kind == 'f':
        return data.astype(numpy.float64, copy=False)
    if data.dtype.kind == 'u':
        return data.astype(numpy.float64) / numpy.iinfo(data.dtype).max
    if data.dtype.kind == 'b':
        return data.astype(numpy.float64)
    raise TypeError("image dtype {} is not supported".format(data.dtype))


def _device_image(data, scaled):
    """(fp64 CUDA tensor (Nw, Nh), device)."""
    torch = _torch()
    if isinstance(data, torch.Tensor):
        if data.dim() != 2:
            raise ValueError("a 2-D image is required; shape {} was given".format(tuple(data.shape)))
        if not data.is_floating_point():
            raise TypeError("device images must be floating point")
        return data.to(device=data.device if data.is_cuda else "cuda", dtype=torch.float64).contiguous()
    host = _as_float_image(data) if scaled else numpy.asarray(data, dtype=numpy.float64)
    if host.ndim != 2:
        raise ValueError("a 2-D image is required; shape {} was given".format(host.shape))
    return torch.from_numpy(numpy.ascontiguousarray(host)).to("cuda")


def _stream(torch, device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def gaussian_half_kernels(sigma, truncate=4.0):
    """Right halves (taps 0..radius) of the kernels ``scipy.ndimage.gaussian_filter1d`` applies for
    ``order=0`` and ``order=2``: the normalised Gaussian, and the Gaussian times the polynomial of
    its second derivative, ``(x^2 / sigma^4 - 1 / sigma^2)``, built by the same recurrence."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = numpy.arange(-radius, radius + 1)
    phi = numpy.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    exponents = numpy.arange(3)
    q = numpy.zeros(3)
    q[0] = 1
    step = numpy.diag(exponents[1:], 1) + numpy.diag(numpy.ones(2) / -sigma2, -1)    # d/dx of q(x) exp(-x^2 / 2 sigma^2)
    for _ in range(2):
        q = step.dot(q)
    second = (x[:, None] ** exponents).dot(q) * phi
    return radius, phi[radius:], second[radius:]


def log_scale_space(image, sigmas):
    """The cube ``[-gaussian_laplace(image, s) * s ** 2 for s in sigmas]`` as a CUDA tensor
    ``(len(sigmas), Nw, Nh)`` (scikit-image stacks the scales on the last axis instead)."""
    torch = _torch()
    lib = _native.load()
    device = image.device
    n_w, n_h = image.shape
    kernels = [gaussian_half_kernels(s) for s in sigmas]
    radii = numpy.array([k[0] for k in kernels], dtype=numpy.int32)
    limit = lib.scb_log_max_radius()
    if radii.max() > limit:
        raise ValueError("max_sigma={} needs a kernel radius of {}; the device tile holds {}".format(
            max(sigmas), radii.max(), limit))
    pitch = int(radii.max()) + 1
    weights = numpy.zeros((len(sigmas), 2, pitch))
    for k, (radius, g0, g2) in enumerate(kernels):
        weights[k, 0, : radius + 1] = g0
        weights[k, 1, : radius + 1] = g2
    sigma2 = numpy.array([float(s) ** 2 for s in sigmas])
    d_radius = torch.from_numpy(radii).to(device)
    d_weights = torch.from_numpy(weights).to(device)
    d_sigma2 = torch.from_numpy(sigma2).to(device)
    cube = torch.empty((len(sigmas), n_w, n_h), dtype=torch.float64, device=device)
    per_scale = lib.scb_log_workspace_bytes(n_w, n_h, 1)
    chunk = max(1, min(len(sigmas), SCALE_SPACE_BYTES // per_scale))
    work = torch.empty(chunk * per_scale, dtype=torch.uint8, device=device)
    for first in range(0, len(sigmas), chunk):
        count = min(chunk, len(sigmas) - first)
        _native.check(lib.scb_log_scale_space(
            n_w, n_h, count, _native.ptr(image), _native.ptr(d_radius[first:]), int(radii[first: first + count].max()),
            _native.ptr(d_weights[first:]), pitch, _native.ptr(d_sigma2[first:]), _native.ptr(cube[first:]),
            _native.ptr(work), work.numel(), _stream(torch, device)), "scb_log_scale_space")
    return cube


def scale_space_peaks(cube, threshold):
    """``peak_local_max(cube, threshold_abs=threshold, footprint=ones((3, 3, 3)),
    exclude_border=False)`` as an ``(n, 3)`` integer array ``(i, j, scale index)``, strongest first
    (ties in C order of ``(i, j, scale)``, as the stable sort of scikit-image leaves them)."""
    torch = _torch()
    lib = _native.load()
    n_sigma, n_w, n_h = cube.shape
    capacity = 1 << 16
    while True:
        peaks = torch.empty((capacity, 3), dtype=torch.int32, device=cube.device)
        values = torch.empty(capacity, dtype=torch.float64, device=cube.device)
        count = torch.zeros(1, dtype=torch.int64, device=cube.device)
        _native.check(lib.scb_log_peaks(n_w, n_h, n_sigma, _native.ptr(cube), float(threshold), _native.ptr(peaks),
                                        _native.ptr(values), capacity, _native.ptr(count),
                                        _stream(torch, cube.device)), "scb_log_peaks")
        found = int(count.item())
        if found == cube.numel() and found > 1:
            return numpy.empty((0, 3), dtype=numpy.int64)        # flat cube: "no peak for a trivial image"
        if found <= capacity:
            break
        capacity = found
    peaks = peaks[:found].cpu().numpy().astype(numpy.int64)
    values = values[:found].cpu().numpy()
    linear = (peaks[:, 0] * n_h + peaks[:, 1]) * n_sigma + peaks[:, 2]
    order = numpy.lexsort((linear, -values))
    return peaks[order]


def _disk_overlap(d, r1, r2):
    ratio1 = min(max((d ** 2 + r1 ** 2 - r2 ** 2) / (2 * d * r1), -1.0), 1.0)
    ratio2 = min(max((d ** 2 + r2 ** 2 - r1 ** 2) / (2 * d * r2), -1.0), 1.0)
    a, b, c, e = -d + r2 + r1, d - r2 + r1, d + r2 - r1, d + r2 + r1
    area = r1 ** 2 * math.acos(ratio1) + r2 ** 2 * math.acos(ratio2) - 0.5 * math.sqrt(abs(a * b * c * e))
    return area / (math.pi * min(r1, r2) ** 2)


def _overlap(blob1, blob2):
    """Fraction of the smaller disc (radius sqrt(2) sigma) covered by the other one."""
    s1, s2 = blob1[2], blob2[2]
    if s1 == 0 and s2 == 0:
        return 0.0
    r1, r2 = (1.0, s2 / s1) if s1 > s2 else (s1 / s2, 1.0)
    unit = max(s1, s2) * math.sqrt(2)                 # distances in units of the larger radius
    d = float(numpy.sqrt(numpy.sum((blob2[:2] / unit - blob1[:2] / unit) ** 2)))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    return _disk_overlap(d, r1, r2)


def prune_blobs(blobs, overlap):
    """scikit-image's ``_prune_blobs``: of two blobs whose discs overlap by more than ``overlap``
    the smaller is dropped.  Pairs are visited in ``cKDTree.query_pairs`` order and a dropped blob
    keeps taking part with radius 0, as there."""
    import scipy.spatial
    reach = 2 * blobs[:, 2].max() * math.sqrt(2)
    pairs = list(scipy.spatial.cKDTree(blobs[:, :2]).query_pairs(reach))
    for i, j in pairs:
        if _overlap(blobs[i], blobs[j]) > overlap:
            if blobs[i, 2] > blobs[j, 2]:
                blobs[j, 2] = 0
            else:
                blobs[i, 2] = 0
    return blobs[blobs[:, 2] > 0]


def blob_detection(data, min_sigma=1, max_sigma=50, num_sigma=10, threshold=0.2, overlap=0.5):
    """Finds blobs in the given image (Laplacian of Gaussian; ``skimage.feature.blob_log``).

    Args:
        data (ndarray or CUDA tensor): An image data.
        min_sigma (float, optional): The minimum standard deviation. Defaults to 1.
        max_sigma (float, optional): The maximum standard deviation. Defaults to 50.
        num_sigma (int, optional): The number of values between `min_sigma` and `max_sigma`.
        threshold (float, optional): The absolute lower bound for scale space maxima.
        overlap (float, optional): A value between 0 and 1.

    Returns:
        ndarray: Blobs detected.  Each row is `(x, y, r)` with `r = sqrt(2) * sigma`.
    """
    image = _device_image(data, scaled=True)
    sigmas = numpy.linspace(0, 1, num_sigma) * (float(max_sigma) - float(min_sigma)) + float(min_sigma)
    cube = log_scale_space(image, sigmas)
    peaks = scale_space_peaks(cube, threshold)
    if len(peaks) == 0:
        blobs = numpy.empty((0, 3))
    else:
        blobs = peaks.astype(numpy.float64)
        blobs[:, 2] = sigmas[peaks[:, 2]]
        blobs = prune_blobs(blobs, overlap)
    blobs[:, 2] = blobs[:, 2] * numpy.sqrt(2)
    _log.info('{} blob(s) were detected'.format(len(blobs)))
    return blobs


def spot_detection(data, roi_size=6, blobs=None, processes=None, **kwargs):
    """Finds spots in the given image.

    Args:
        data (ndarray or CUDA tensor): An image data.
        roi_size (float, optional): A half of the ROI size. Defaults to 6.
        blobs (ndarray, optional): Blobs. Defaults to `None`. See also `blob_detection`.
        processes: accepted for compatibility; every blob is fitted in one kernel launch.

    Returns:
        ndarray: Spots detected.  Each row is
            `(center_x, center_y, intensity, bg, height, sigma)`.
    """
    torch = _torch()
    lib = _native.load()
    if blobs is None:
        blobs = blob_detection(data, **kwargs)
    blobs = numpy.ascontiguousarray(numpy.asarray(blobs, dtype=numpy.float64))
    if blobs.size == 0:
        _log.info('0 spot(s) were detected')
        return numpy.array([])
    if blobs.ndim != 2 or blobs.shape[1] < 2:
        raise ValueError("blobs must have rows (x, y[, r]); shape {} was given".format(blobs.shape))
    image = _device_image(data, scaled=False)
    n_w, n_h = image.shape
    d_blobs = torch.from_numpy(blobs).to(image.device)
    spots = torch.zeros((len(blobs), 6), dtype=torch.float64, device=image.device)
    status = torch.full((len(blobs),), -1, dtype=torch.int32, device=image.device)
    _native.check(lib.scb_spot_fit(n_w, n_h, _native.ptr(image), len(blobs), _native.ptr(d_blobs), blobs.shape[1],
                                   float(roi_size), MAX_FIT_ITERATIONS, _native.ptr(spots), _native.ptr(status),
                                   _stream(torch, image.device)), "scb_spot_fit")
    status = status.cpu().numpy()
    spots = spots.cpu().numpy()
    for code in status[status != 0]:
        _log.debug(_SKIP_REASONS.get(int(code), "spot_detection skip a blob."))
    spots = spots[status == 0]
    _log.info('{} spot(s) were detected'.format(len(spots)))
    return spots
