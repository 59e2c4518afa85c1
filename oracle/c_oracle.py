"""ctypes wrapper of ``oracle/overlay_oracle.c`` (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboverlay_oracle.so")


class Geometry(ctypes.Structure):
    _fields_ = [("n_w", ctypes.c_int), ("n_h", ctypes.c_int), ("n_radial", ctypes.c_int),
                ("n_depth_keys", ctypes.c_int), ("pixel_length", ctypes.c_double),
                ("resolution", ctypes.c_double), ("depth_cutoff", ctypes.c_double),
                ("focal", ctypes.c_double * 3)]


def build(force=False):
    src = os.path.join(HERE, "overlay_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_table_scale.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def geometry(params):
    """``orc_geometry`` from an oracle parameter dict."""
    n_radial = len(numpy.arange(0.0, params["radial_cutoff"], 1e-9))
    g = Geometry(n_w=params["image_size"][0], n_h=params["image_size"][1], n_radial=n_radial,
                 n_depth_keys=int((params["depth_cutoff"] + 1e-9) / 1e-9) + 1,
                 pixel_length=params["pixel_length"] / params["magnification"], resolution=1e-9,
                 depth_cutoff=params["depth_cutoff"])
    for i, v in enumerate(params["focal_point"]):
        g.focal[i] = v
    return g


def table_from_radial(prof):
    prof = numpy.ascontiguousarray(prof, dtype=numpy.float64)
    side = 2 * (len(prof) - 1) + 1
    T = numpy.empty((side, side))
    lib().orc_table_from_radial(_p(prof), ctypes.c_int(len(prof)), _p(T))
    return T


def sat_from_table(T):
    """(S int64 (side+1, side+1), inv_scale) with the device's quantisation rule."""
    T = numpy.ascontiguousarray(T, dtype=numpy.float64)
    side = T.shape[0]
    scale = lib().orc_table_scale(_p(T), ctypes.c_int(side))
    S = numpy.empty((side + 1, side + 1), dtype=numpy.int64)
    lib().orc_sat_int64(_p(T), ctypes.c_int(side), ctypes.c_double(scale), _p(S))
    return S, 1.0 / scale


def _spots(depth, x, y, weight):
    return [numpy.ascontiguousarray(a, dtype=numpy.float64) for a in (depth, x, y, weight)]


def render_sat(g, depth, x, y, weight, sats, inv_scale, slot_of_key):
    depth, x, y, weight = _spots(depth, x, y, weight)
    sats = numpy.ascontiguousarray(sats, dtype=numpy.int64)
    inv_scale = numpy.ascontiguousarray(inv_scale, dtype=numpy.float64)
    slot_of_key = numpy.ascontiguousarray(slot_of_key, dtype=numpy.int32)
    out = numpy.zeros((g.n_w, g.n_h))
    missing = lib().orc_render_sat(ctypes.byref(g), ctypes.c_int64(len(x)), _p(depth), _p(x), _p(y), _p(weight),
                                   _p(sats), _p(inv_scale), _p(slot_of_key), _p(out))
    assert missing == 0, "{} spots without a table".format(missing)
    return out


def render_bruteforce(g, depth, x, y, weight, tables, slot_of_key, n_threads=1):
    depth, x, y, weight = _spots(depth, x, y, weight)
    tables = numpy.ascontiguousarray(tables, dtype=numpy.float64)
    slot_of_key = numpy.ascontiguousarray(slot_of_key, dtype=numpy.int32)
    out = numpy.zeros((g.n_w, g.n_h))
    missing = lib().orc_render_bruteforce(ctypes.byref(g), ctypes.c_int64(len(x)), _p(depth), _p(x), _p(y),
                                          _p(weight), _p(tables), _p(slot_of_key), _p(out), ctypes.c_int(n_threads))
    assert missing == 0
    return out


def detector_frame(photons, qe, background, is_cmos, rn_values, rn_weights, readout_sigma, fullwell, adc0, bit,
                   seed, n_threads=0):
    photons = numpy.ascontiguousarray(photons, dtype=numpy.float64)
    vals = numpy.ascontiguousarray(rn_values if rn_values is not None else [0.0], dtype=numpy.float64)
    w = numpy.asarray(rn_weights if rn_weights is not None else [1.0], dtype=numpy.float64)
    cdf = numpy.ascontiguousarray(numpy.cumsum(w / w.sum()))
    out = numpy.empty(photons.shape)
    lib().orc_detector_frame(ctypes.c_int64(photons.size), _p(photons), ctypes.c_double(qe), ctypes.c_double(background),
                             ctypes.c_int(int(is_cmos)), _p(vals), _p(cdf), ctypes.c_int(len(vals)),
                             ctypes.c_double(readout_sigma), ctypes.c_double(fullwell), ctypes.c_double(adc0),
                             ctypes.c_int(bit), ctypes.c_uint64(seed), _p(out), ctypes.c_int(n_threads))
    return out


def move_points(coords, sigma, seed):
    coords = numpy.ascontiguousarray(coords, dtype=numpy.float64)
    sigma = numpy.ascontiguousarray(sigma, dtype=numpy.float64)
    lib().orc_move_points(ctypes.c_int64(coords.shape[0]), ctypes.c_int(coords.shape[1]), _p(coords), _p(sigma),
                          ctypes.c_uint64(seed))
    return coords
