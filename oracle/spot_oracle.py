"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the reference's spot detector,
``/root/reference/src/scopyon/analysis/spot_detection.py``:

  * ``spot_detection``  (``:139-174``)  -> ``spot_detection`` here, built from
    ``mean_background`` (``:49-58``), ``planar_background`` (``:60-76``),
    ``background`` (``:78-82``), ``weighted_com`` (``:88-93``), ``fitgaussian``
    (``:95-103``) and the per-blob driver ``__spot_detection`` (``:110-137``).
    Same library calls as the reference (``scipy.optimize.least_squares`` with its
    defaults), so this half is PINNED: ``tests/test_oracle_vs_reference.py`` runs
    it against the live reference with the blobs passed in, and
    ``tests/golden/spots_case.npz`` carries the reference's output.

  * ``blob_detection``  (``:14-47``) calls ``skimage.feature.blob_log``.
    scikit-image is an optional dependency of the reference (``:37-40``: it raises
    ``ImportError`` when absent), is not listed in ``uv.lock`` and is not installed in
    this image, so there is no pinned version to name.  ``blob_log`` below restates the
    published algorithm of scikit-image 0.19-0.25 (``skimage/feature/blob.py``:
    ``blob_log``, ``_prune_blobs``, ``_blob_overlap``, ``_compute_disk_overlap``;
    ``skimage/feature/peak.py``: ``peak_local_max`` with a 3x3x3 footprint,
    ``threshold_abs`` and ``exclude_border=False``) on top of the SAME scipy calls
    scikit-image makes (``scipy.ndimage.gaussian_laplace``, ``maximum_filter``,
    ``scipy.spatial.cKDTree.query_pairs``).  PARITY UNPINNED for this half: no
    scikit-image here to run it against, and the reference holds no golden vectors
    for it.
"""
import math

import numpy
import scipy.ndimage
import scipy.optimize
import scipy.spatial

__all__ = ["sigma_list", "log_cube", "peak_local_max_3d", "prune_blobs", "blob_log", "blob_detection",
           "spot_detection", "fit_blob"]


# --------------------------------------------------------------------------------------
# blob_log (scikit-image) -- restated
def sigma_list(min_sigma, max_sigma, num_sigma):
    """``blob_log``: ``np.linspace(0, 1, num_sigma)[:, None] * (max - min) + min`` (log_scale=False)."""
    scale = numpy.linspace(0, 1, num_sigma)
    return scale * (float(max_sigma) - float(min_sigma)) + float(min_sigma)


def log_cube(image, sigmas):
    """``[-gaussian_laplace(image, s) * mean(s) ** 2 for s in sigma_list]`` stacked on the LAST axis."""
    image = numpy.asarray(image, dtype=numpy.float64)
    planes = [-scipy.ndimage.gaussian_laplace(image, [s, s]) * s ** 2 for s in sigmas]
    return numpy.stack(planes, axis=-1)


def peak_local_max_3d(cube, threshold):
    """``peak_local_max(cube, threshold_abs=threshold, footprint=ones((3, 3, 3)),
    exclude_border=False)``: voxels equal to the maximum of their 3x3x3 neighbourhood (edges
    replicated, ``mode='nearest'``) and ``> threshold``; none when every voxel qualifies (flat
    image); ordered by decreasing value, ties in C order (stable sort)."""
    if cube.size == 1:
        mask = cube > threshold
    else:
        peak = scipy.ndimage.maximum_filter(cube, footprint=numpy.ones((3, 3, 3)), mode='nearest')
        mask = cube == peak
        if numpy.all(mask):
            mask[:] = False
        mask &= cube > threshold
    coord = numpy.nonzero(mask)
    order = numpy.argsort(-cube[coord], kind="stable")
    return numpy.transpose(coord)[order]


def _disk_overlap(d, r1, r2):
    ratio1 = (d ** 2 + r1 ** 2 - r2 ** 2) / (2 * d * r1)
    ratio1 = min(max(ratio1, -1.0), 1.0)
    acos1 = math.acos(ratio1)
    ratio2 = (d ** 2 + r2 ** 2 - r1 ** 2) / (2 * d * r2)
    ratio2 = min(max(ratio2, -1.0), 1.0)
    acos2 = math.acos(ratio2)
    a = -d + r2 + r1
    b = d - r2 + r1
    c = d + r2 - r1
    e = d + r2 + r1
    area = r1 ** 2 * acos1 + r2 ** 2 * acos2 - 0.5 * math.sqrt(abs(a * b * c * e))
    return area / (math.pi * (min(r1, r2) ** 2))


def _blob_overlap(blob1, blob2):
    root_ndim = math.sqrt(2)
    if blob1[-1] == blob2[-1] == 0:
        return 0.0
    elif blob1[-1] > blob2[-1]:
        max_sigma = blob1[-1]
        r1, r2 = 1.0, blob2[-1] / blob1[-1]
    else:
        max_sigma = blob2[-1]
        r2, r1 = 1.0, blob1[-1] / blob2[-1]
    pos1 = blob1[:2] / (max_sigma * root_ndim)
    pos2 = blob2[:2] / (max_sigma * root_ndim)
    d = float(numpy.sqrt(numpy.sum((pos2 - pos1) ** 2)))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    return _disk_overlap(d, r1, r2)


def prune_blobs(blobs, overlap):
    """``_prune_blobs(blobs, overlap, sigma_dim=1)``: for every pair closer than
    ``2 * max sigma * sqrt(2)`` whose discs overlap by more than ``overlap``, the smaller one is
    dropped (its sigma zeroed in place, which later pairs see)."""
    blobs = numpy.array(blobs, dtype=numpy.float64)
    sigma = blobs[:, -1].max()
    distance = 2 * sigma * math.sqrt(blobs.shape[1] - 1)
    tree = scipy.spatial.cKDTree(blobs[:, :-1])
    pairs = numpy.array(list(tree.query_pairs(distance)))
    if len(pairs) == 0:
        return blobs
    for (i, j) in pairs:
        blob1, blob2 = blobs[i], blobs[j]
        if _blob_overlap(blob1, blob2) > overlap:
            if blob1[-1] > blob2[-1]:
                blob2[-1] = 0
            else:
                blob1[-1] = 0
    return numpy.stack([b for b in blobs if b[-1] > 0])


def blob_log(image, min_sigma=1, max_sigma=50, num_sigma=10, threshold=0.2, overlap=0.5):
    sigmas = sigma_list(min_sigma, max_sigma, num_sigma)
    cube = log_cube(image, sigmas)
    peaks = peak_local_max_3d(cube, threshold)
    if peaks.size == 0:
        return numpy.empty((0, 3))
    lm = peaks.astype(numpy.float64)
    lm[:, -1] = sigmas[peaks[:, -1]]
    return prune_blobs(lm, overlap)


def blob_detection(data, min_sigma=1, max_sigma=50, num_sigma=10, threshold=0.2, overlap=0.5):
    """spot_detection.py:14-47: ``blob_log`` then the radius column times sqrt(2)."""
    blobs = blob_log(data, min_sigma=min_sigma, max_sigma=max_sigma, num_sigma=num_sigma,
                     threshold=threshold, overlap=overlap)
    blobs[:, 2] = blobs[:, 2] * numpy.sqrt(2)
    return blobs


# --------------------------------------------------------------------------------------
# spot_detection.py:49-137 -- per-blob background plane and Gaussian fit
def _mean_background(roi):
    m, n = roi.shape
    left = roi[0, : -1].sum()
    right = roi[-1, : -1].sum()
    bottom = roi[: -1, 0].sum()
    top = roi[: -1, -1].sum()
    tot = left + right + bottom + top - (roi[0, 0] + roi[0, -1] + roi[-1, 0] + roi[-1, -1])
    return tot / (2 * (m + n - 2))


def _planar_background(roi):
    m, n = roi.shape
    rows, cols = numpy.arange(m), numpy.arange(n)

    def residuals(p):
        a5, a6, a7 = p
        return numpy.concatenate([
            roi[0, :] - (cols * a6 + a7),
            roi[-1, :] - ((m - 1) * a5 + cols * a6 + a7),
            roi[:, 0] - (rows * a5 + a7),
            roi[:, -1] - (rows * a5 + (n - 1) * a6 + a7)])
    res = scipy.optimize.least_squares(residuals, (0.0, 0.0, _mean_background(roi)))
    return None if res.success <= 0 else res.x


def _gaussian(p, X, Y):
    a1, a2, a3, a4 = p
    return a1 * numpy.exp(-((X - a2) ** 2 + (Y - a3) ** 2) / a4)


def _fit_gaussian(data, roi_size):
    X, Y = numpy.indices(data.shape)
    total = data.sum()
    start = (255, (X * data).sum() / total, (Y * data).sum() / total, roi_size / 2)
    res = scipy.optimize.least_squares(lambda p: numpy.ravel(_gaussian(p, X, Y) - data), start)
    return None if res.success <= 0 else res.x


def fit_blob(blob, data, roi_size):
    """spot_detection.py:110-137.  Returns the 6-tuple or None when the reference skips the blob."""
    x, y = blob[0], blob[1]
    x0, x1 = int(x - roi_size), int(x + roi_size) + 1
    y0, y1 = int(y - roi_size), int(y + roi_size) + 1
    x0, x1 = max(0, x0), min(data.shape[0], x1)
    y0, y1 = max(0, y0), min(data.shape[1], y1)
    roi = data[x0: x1, y0: y1]
    if roi.sum() <= 0:
        return None
    plane = _planar_background(roi)
    if plane is None:
        return None
    m, n = roi.shape
    bg = numpy.arange(m)[:, None] * plane[0] + numpy.arange(n)[None, :] * plane[1] + plane[2]
    res = _fit_gaussian(roi.astype(numpy.float64) - bg, roi_size)
    if res is None:
        return None
    height, cx, cy, sigma = res
    if not (0 <= cx < m and 0 <= cy < n):
        return None
    X, Y = numpy.indices(roi.shape)
    intensity = _gaussian(res, X, Y).sum()
    return (cx + x0, cy + y0, intensity, bg.sum(), height, sigma)


def spot_detection(data, roi_size=6, blobs=None, **kwargs):
    """spot_detection.py:139-174: rows ``(center_x, center_y, intensity, bg, height, sigma)``."""
    if blobs is None:
        blobs = blob_detection(data, **kwargs)
    spots = [s for s in (fit_blob(b, data, roi_size) for b in blobs) if s is not None]
    return numpy.array(spots)

