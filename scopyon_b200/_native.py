"""ctypes binding of ``libscopyon_b200.so`` (the C ABI declared in ``include/scopyon_b200.h``).

There is no CPU fallback: if the library is missing the import of any compute
entry point raises, and every non-zero status from the library raises ``NativeError``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscopyon_b200.so")

c_f64p = ctypes.c_void_p   # device pointers travel as integers (tensor.data_ptr())
c_ptr = ctypes.c_void_p

PSF_BORN_WOLF, PSF_GAUSSIAN = 0, 1
DET_CMOS, DET_EMCCD, DET_CCD = 0, 1, 2
FPN_NONE, FPN_PIXEL, FPN_COLUMN = 0, 1, 2
F32, F64 = 0, 1

DETECTOR_CODES = {"CMOS": DET_CMOS, "EMCCD": DET_EMCCD, "CCD": DET_CCD}
FPN_CODES = {"none": FPN_NONE, "pixel": FPN_PIXEL, "column": FPN_COLUMN}


class NativeError(RuntimeError):
    pass


class Geometry(ctypes.Structure):
    _fields_ = [
        ("n_w", ctypes.c_int32), ("n_h", ctypes.c_int32),
        ("n_radial", ctypes.c_int32), ("n_depth_keys", ctypes.c_int32),
        ("sat_modulus", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("pixel_length", ctypes.c_double), ("resolution", ctypes.c_double),
        ("depth_cutoff", ctypes.c_double), ("focal", ctypes.c_double * 3),
        ("box_peak", ctypes.c_double),
    ]


class Photophysics(ctypes.Structure):
    _fields_ = [
        ("amplitude0", ctypes.c_double), ("penetration_depth", ctypes.c_double),
        ("x_sec", ctypes.c_double), ("quantum_yield", ctypes.c_double),
        ("absorb_frac", ctypes.c_double), ("norm_scale", ctypes.c_double),
        ("budget_scale", ctypes.c_double),
    ]


class Detector(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32), ("fpn_type", ctypes.c_int32),
        ("bit", ctypes.c_int32), ("background_on", ctypes.c_int32),
        ("qe", ctypes.c_double), ("background", ctypes.c_double),
        ("readout_noise", ctypes.c_double), ("emgain", ctypes.c_double),
        ("fullwell", ctypes.c_double), ("adc_offset", ctypes.c_double),
        ("fpn_count", ctypes.c_double),
    ]


# symbol -> (restype, argtypes); must list every function include/scopyon_b200.h declares
_Vec3 = ctypes.c_double * 3
SIGNATURES = {
    "scb_version": (ctypes.c_int, []),
    "scb_last_error": (ctypes.c_char_p, []),
    "scb_psf_radial_build": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr]),
    "scb_psf_sat_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "scb_psf_sat_table_entries": (ctypes.c_int64, [ctypes.c_int, ctypes.c_int]),
    "scb_psf_sat_slots": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "scb_psf_sat_build": (ctypes.c_int, [
        c_ptr, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, ctypes.c_int, c_ptr, c_ptr, ctypes.c_size_t,
        c_ptr]),
    "scb_diffuse": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.POINTER(ctypes.c_double),
        c_ptr, c_ptr, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), c_ptr]),
    "scb_transition_states": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, c_ptr, c_ptr, ctypes.c_int, c_ptr]),
    "scb_place_uniform": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, c_ptr, c_ptr, c_ptr,
        ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), c_ptr]),
    "scb_emit_bleach": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
        ctypes.c_double, ctypes.c_double, ctypes.POINTER(Photophysics), c_ptr, c_ptr, c_ptr, c_ptr]),
    "scb_replay_frames": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
        c_ptr, c_ptr, c_ptr, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double,
        ctypes.POINTER(Photophysics), c_ptr, c_ptr]),
    "scb_render_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(Geometry), ctypes.c_int64]),
    "scb_render_expected": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, c_ptr,
        c_ptr, ctypes.c_int, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_emit_bleach_rows": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_int64, c_ptr, c_ptr, ctypes.c_double, ctypes.c_double,
        ctypes.POINTER(Photophysics), c_ptr, c_ptr, c_ptr, c_ptr]),
    "scb_render_expected_rows": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, c_ptr, c_ptr,
        ctypes.c_int, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_movie_frames": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
        c_ptr, c_ptr, c_ptr, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double,
        ctypes.POINTER(Photophysics), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "scb_render_frames_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(Geometry), ctypes.c_int64, ctypes.c_int]),
    "scb_render_expected_frames": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_render_expected_rows_frames": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_render_expected_frames_ordered": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_render_expected_frames_planned": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, ctypes.c_size_t, c_ptr, ctypes.c_int, c_ptr]),
    "scb_gaussian_tc_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(Geometry), ctypes.c_int64]),
    "scb_render_gaussian_tc": (ctypes.c_int, [
        ctypes.POINTER(Geometry), ctypes.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_int, ctypes.c_int,
        c_ptr, ctypes.c_size_t, c_ptr, c_ptr]),
    "scb_profile_begin": (ctypes.c_int, [ctypes.c_int]),
    "scb_profile_end": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "scb_adc_offsets": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_int64, ctypes.c_double, ctypes.c_double, c_ptr, ctypes.c_int, c_ptr]),
    "scb_detector_adc": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(Detector), ctypes.c_int32, ctypes.c_int32,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
        c_ptr, ctypes.c_size_t, c_ptr]),
    "scb_detector_adc_frames": (ctypes.c_int, [
        ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(Detector), ctypes.c_int32, ctypes.c_int32,
        ctypes.c_int, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr, c_ptr, ctypes.c_size_t, c_ptr]),
    "scb_detector_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int32]),
    "scb_frames_minmax": (ctypes.c_int, [c_ptr, ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr, c_ptr]),
    "scb_frames_to_8bit": (ctypes.c_int, [
        c_ptr, ctypes.c_int64, ctypes.c_int, c_ptr, ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.c_double, c_ptr, c_ptr]),
    "scb_log_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "scb_log_max_radius": (ctypes.c_int, []),
    "scb_log_scale_space": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, ctypes.c_int, c_ptr, ctypes.c_int, c_ptr, c_ptr,
        c_ptr, ctypes.c_size_t, c_ptr]),
    "scb_log_peaks": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ptr, ctypes.c_double, c_ptr, c_ptr, ctypes.c_int64, c_ptr, c_ptr]),
    "scb_spot_fit": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_int, c_ptr, ctypes.c_int64, c_ptr, ctypes.c_int, ctypes.c_double, ctypes.c_int, c_ptr,
        c_ptr, c_ptr]),
    "scb_host_widen_start": (ctypes.c_int64, [c_ptr, c_ptr, ctypes.c_int64, c_ptr, ctypes.c_int]),
    "scb_host_widen_wait": (ctypes.c_int, [ctypes.c_int64]),
    "scb_host_widen_threads": (ctypes.c_int, [ctypes.c_int]),
    "scb_host_widen_affinity": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "scb_host_bandwidth": (ctypes.c_int, [
        ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "scb_frames_to_u16": (ctypes.c_int, [c_ptr, ctypes.c_int64, ctypes.c_int, c_ptr, c_ptr]),
    "scb_device_sm_count": (ctypes.c_int, []),
    # known-answer hooks for the tests
    "scb_philox4x32_10": (None, [ctypes.POINTER(ctypes.c_uint32)] * 3),
    "scb_test_poisson_inversion": (ctypes.c_int, [ctypes.c_int64, c_ptr, c_ptr, c_ptr, c_ptr]),
}

_EXTRA = {}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "{} is missing: build it with `python -m scopyon_b200.build` "
            "(scopyon_b200 has no CPU fallback)".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in list(SIGNATURES.items()) + list(_EXTRA.items()):
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.scb_version() != 100:
        raise NativeError("libscopyon_b200.so version mismatch: {}".format(lib.scb_version()))
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().scb_last_error().decode("utf-8", "replace")
        raise NativeError("{} failed with status {}: {}".format(what, status, msg))


def vec3(values):
    return _Vec3(*[float(v) for v in values])


def ptr(tensor):
    """Device (or NULL) pointer of a torch tensor."""
    return None if tensor is None else ctypes.c_void_p(tensor.data_ptr())
