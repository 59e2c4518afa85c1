"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's image-formation hot path
(``/root/reference/src/scopyon/_epifm.py`` + ``sampling.py``).  It exists to
CHECK the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``scopyon_b200/`` imports this module, and the product has no
CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs every
function below against the live, unmodified reference in the build container
(via ``oracle/ref_shim.py``), and ``tests/golden/*.npz`` (made by
``oracle/make_golden.py`` from the live reference) pin it on machines where
``/root/reference`` is absent.  The six integrals the reference's own test
prints (``test/test_epifm.py:30-42,58-70``) are asserted in
``tests/test_oracle_golden.py``.

All inputs are plain SI floats / numpy arrays (a ``dict`` of parameters, see
``PARAM_KEYS``); no reference object is needed at run time.
"""
import math

import numpy

# physical constants, reference ``constants.py:12-17`` (CODATA-2018 exact values)
N_A = 6.02214076e+23
HC = 6.62607015e-34 * 299792458

RES = 1e-9  # table resolution, ``_epifm.py:59-60``

PARAM_KEYS = (
    "image_size", "pixel_length", "magnification", "focal_point",
    "psf_type", "psf_wavelength", "psf_radial_width", "radial_cutoff", "depth_cutoff",
    "psf_normalization", "fluoem_norm_sum",
    "source_flux_density", "source_wavelength", "source_angle",
    "quantum_yield", "abs_coefficient", "fluorophore_radius",
    "bleaching_switch", "bleaching_half_life", "background_switch", "background_mean",
    "detector_type", "QE", "readout_noise", "emgain", "exposure_time",
    "adc_bit", "adc_offset", "adc_fullwell", "fpn_type", "fpn_count",
    "shutter_switch", "shutter_start_time", "shutter_end_time",
)


# --------------------------------------------------------------------------- PSF
def radial_grid(radial_cutoff):
    """``_epifm.py:91``: r = arange(0, cutoff, 1 nm)."""
    return numpy.arange(0.0, radial_cutoff, RES, dtype=float)


def born_wolf_radial(r, z, wave_length):
    """``_epifm.py:181-213`` (__get_born_wolf_distribution_old), NA = 1.4, 100 rho terms."""
    NA = 1.4
    k = 2.0 * numpy.pi / wave_length
    alpha = k * NA
    gamma = k * numpy.power(NA / 2, 2)
    N = 100
    drho = 1.0 / N
    rho = numpy.arange(1, N + 1) * drho
    from scipy.special import j0
    J0 = j0(r[:, None] * alpha * rho[None, :])
    Y = numpy.exp(-2 * 1.j * z * gamma * rho * rho) * rho * drho
    I_sum = (Y * J0).sum(axis=1)
    # scalar abs()**2 per element like the reference (numpy's scalar power differs from
    # the array square in the last bit)
    psf = numpy.array([abs(x) ** 2 for x in I_sum])
    psf *= alpha * alpha / numpy.pi
    return psf


def gaussian_radial(r, radial_width):
    """``_epifm.py:133-134``."""
    return numpy.exp(-0.5 * (r / radial_width) ** 2) / (2 * numpy.pi * radial_width * radial_width)


def radial_profile(params, depth):
    r = radial_grid(params["radial_cutoff"])
    if params["psf_type"] == "Gaussian":
        return gaussian_radial(r, params["psf_radial_width"])
    return born_wolf_radial(r, depth, params["psf_wavelength"])


def radial_to_cartesian(radial, radial_distribution):
    """``_epifm.py:98-126``.  The reference builds ``scipy.interpolate.interp1d(radial,
    dist)`` (kind='linear'); for 1-D float64 data without extrapolation scipy delegates
    that to ``numpy.interp``, which is called here directly (bit-identical, checked
    against the live reference in tests)."""
    n = radial.size
    side = 2 * (n - 1) + 1
    X, Y = numpy.meshgrid(numpy.arange(side), numpy.arange(side))
    X = (X.ravel() - (n - 1)) * RES
    Y = (Y.ravel() - (n - 1)) * RES
    R = numpy.sqrt(X ** 2 + Y ** 2)
    R[R > radial.max()] = radial.max()
    P = numpy.interp(R, radial, radial_distribution)
    return P.reshape((side, side))


def depth_key(depth, depth_cutoff):
    """``_epifm.py:76-84``: (key, table depth).  key = -1 freezes the PSF at the cutoff."""
    depth = abs(depth)
    if depth < depth_cutoff + RES:
        key = int(depth / RES)
        return key, key * RES
    return -1, depth_cutoff


class PsfTables:
    """Cache of Cartesian tables by depth key (``_epifm.py:86-88``)."""

    def __init__(self, params):
        self.params = params
        self.radial = radial_grid(params["radial_cutoff"])
        self.tables = {}
        self.sats = {}

    def get(self, depth):
        key, zdepth = depth_key(depth, self.params["depth_cutoff"])
        if key not in self.tables:
            self.tables[key] = radial_to_cartesian(self.radial, radial_profile(self.params, zdepth))
        return key, self.tables[key]

    def sat(self, depth):
        """fp64 summed-area table of the same 1-nm samples (fast oracle mode)."""
        key, table = self.get(depth)
        if key not in self.sats:
            S = numpy.zeros((table.shape[0] + 1, table.shape[1] + 1))
            S[1:, 1:] = table.cumsum(axis=0).cumsum(axis=1)
            self.sats[key] = S
        return key, self.sats[key]


# ------------------------------------------------------------------- overlay (a15)
def overlay_edges(xi, n_pixel, pixel_length, n_table):
    """``_epifm.py:228-253`` for one axis: returns (first pixel index, edge sample
    indices) or (0, None) when the spot touches no pixel edge list.
    ``n_table`` = table.shape[axis] (1999)."""
    signal_width = RES * (n_table - 1)
    expected_width = n_pixel * pixel_length
    imin = math.floor((expected_width * 0.5 + xi - signal_width * 0.5) / pixel_length)
    imax = math.ceil((expected_width * 0.5 + xi + signal_width * 0.5) / pixel_length)
    iarray = numpy.arange(max(0, imin), min(n_pixel, imax) + 1)
    left = (iarray * pixel_length - (expected_width * 0.5 + xi - signal_width * 0.5)) / RES
    left = numpy.ceil(left).astype(int)
    if len(left) == 0:
        return 0, None
    left[0] = max(left[0], 0)
    left[-1] = min(left[-1], n_table)
    return int(iarray[0]), left


def overlay_signal_exact(expected, table, p_i, pixel_length, normalization):
    """``_epifm.py:224-282`` verbatim semantics (slice sums), slow; for small cases."""
    _, xi, yi = p_i
    i_first, left = overlay_edges(xi, expected.shape[0], pixel_length, table.shape[0])
    if left is None:
        return
    j_first, top = overlay_edges(yi, expected.shape[1], pixel_length, table.shape[1])
    if top is None:
        return
    unit_area = RES * RES
    for a in range(len(left) - 1):
        for b in range(len(top) - 1):
            photons = table[left[a]: left[a + 1], top[b]: top[b + 1]].sum() * unit_area
            if photons > 0:
                expected[i_first + a, j_first + b] += photons * normalization


def overlay_signal_sat(expected, sat, p_i, pixel_length, normalization):
    """Same pixel box sums evaluated from an fp64 summed-area table: identical index
    arithmetic, only the summation order differs (|rel diff| ~1e-13)."""
    _, xi, yi = p_i
    n_table = sat.shape[0] - 1
    i_first, left = overlay_edges(xi, expected.shape[0], pixel_length, n_table)
    if left is None or len(left) < 2:
        return
    j_first, top = overlay_edges(yi, expected.shape[1], pixel_length, n_table)
    if top is None or len(top) < 2:
        return
    # python slices with start > stop are empty
    l0, l1 = left[:-1], numpy.maximum(left[1:], left[:-1])
    t0, t1 = top[:-1], numpy.maximum(top[1:], top[:-1])
    box = (sat[numpy.ix_(l1, t1)] - sat[numpy.ix_(l0, t1)]
           - sat[numpy.ix_(l1, t0)] + sat[numpy.ix_(l0, t0)])
    photons = box * (RES * RES)
    photons[photons <= 0] = 0.0
    expected[i_first: i_first + len(l0), j_first: j_first + len(t0)] += photons * normalization


# --------------------------------------------------------------- photophysics (a7-a10)
def snells_law(params):
    """``_epifm.py:1362-1428``: (amplitude, penetration depth)."""
    P_0 = params["source_flux_density"]
    wave_length = params["source_wavelength"]
    E_wl = HC / wave_length
    N_0 = P_0 / E_wl
    A2_Is = N_0
    A2_Ip = N_0
    theta_in = params["source_angle"]
    sin = numpy.sin(theta_in)
    cos = numpy.cos(theta_in)
    sin2 = sin ** 2
    cos2 = cos ** 2
    n_1 = 1.46
    n_2 = 1.384
    r = n_2 / n_1
    r2 = r ** 2
    if sin2 / r2 < 1:
        return N_0, numpy.inf
    A2_x = A2_Ip * (4 * cos2 * (sin2 - r2) / (r2 ** 2 * cos2 + sin2 - r2))
    A2_y = A2_Is * (4 * cos2 / (1 - r2))
    A2_z = A2_Ip * (4 * cos2 * sin2 / (r2 ** 2 * cos2 + sin2 - r2))
    amplitude = ((A2_x + A2_z) + A2_y) / 2
    penetration_depth = wave_length / (4.0 * numpy.pi * numpy.sqrt(n_1 ** 2 * sin2 - n_2 ** 2))
    return amplitude, penetration_depth


def emit_photons(amplitude, unit_time, abs_coeff, QY, fluorophore_radius):
    """``_epifm.py:1343-1360``."""
    x_sec = numpy.log(10) * abs_coeff * 0.1 / N_A
    n_abs = amplitude * x_sec * unit_time
    fluorophore_volume = (4.0 / 3.0) * numpy.pi * numpy.power(fluorophore_radius, 3)
    fluorophore_depth = 2.0 * fluorophore_radius
    A = (abs_coeff * 0.1 / N_A) * (1.0 / fluorophore_volume) * fluorophore_depth
    return QY * n_abs * (1.0 - numpy.power(10.0, -A))


def _emit(params, amplitude, unit_time):
    return emit_photons(amplitude, unit_time, params["abs_coefficient"],
                        params["quantum_yield"], params["fluorophore_radius"])


def photon_budget_scale(params):
    """``_epifm.py:1298-1301,1486-1491``: budget = Exp(1) * beta * N_emit0."""
    amplitude0, _ = snells_law(params)
    N_emit0 = _emit(params, amplitude0, 1.0)
    beta = params["bleaching_half_life"] / numpy.log(2.0)
    return beta, N_emit0


def emitted(params, depth, unit_time):
    """``_epifm.py:1281-1292``: photons emitted by one molecule during one sub-step."""
    amplitude, penet_depth = snells_law(params)
    amplitude = amplitude * numpy.exp(-abs(depth) / penet_depth)
    return _emit(params, amplitude, unit_time)


def spot_weight(params, N_emit, p_state=1.0):
    """``_epifm.py:1308-1315`` (filters off): PSF weight of one (particle, sub-step)."""
    normalization = params["fluoem_norm_sum"] * params["psf_normalization"]
    normalization *= p_state * N_emit / (4.0 * numpy.pi)
    return normalization


# ------------------------------------------------------------- frame assembly (a6,a16)
def frame_windows(times, frame_index, start_time, exposure_time, params):
    """``_epifm.py:1149-1163,1183-1193``: list of (snapshot index, unit_time)."""
    times = numpy.asarray(times, dtype=float)
    t = start_time + exposure_time * frame_index
    if params.get("shutter_switch"):
        t = max(t, params["shutter_start_time"])
        exposure_time = max(0.0, min(t + exposure_time, params["shutter_end_time"]))
    start_index = numpy.searchsorted(times, t, side='right')
    if start_index != 0:
        start_index -= 1
    stop_index = numpy.searchsorted(times, t + exposure_time, side='left')
    out = []
    idx = list(range(start_index, stop_index))
    for i, k in enumerate(idx):
        current_time = times[k] if i != 0 else t
        next_time = times[idx[i + 1]] if i + 1 < len(idx) else t + exposure_time
        unit_time = next_time - current_time
        if unit_time < 1e-13:
            continue
        out.append((k, unit_time))
    return out, exposure_time


def expected_frame(input_data, params, frame_index=0, start_time=0.0, exposure_time=None,
                   fluorescence_states=None, budget_draw=None, psf=None, exact=False):
    """Pre-noise photon image + true_data of one frame.

    Follows ``output_frame`` (``_epifm.py:1121-1216``) + ``get_molecule_plane`` /
    ``__overlay_molecule_plane`` (``:1227-1333``).  ``input_data`` is the formatted
    list of ``(time, (N,5) [depth, x, y, id, p_state])``.  ``budget_draw(m_id)``
    supplies the Exp(1) variate for a molecule's photon budget (the reference
    draws ``rng.exponential`` in particle order, ``:1489``).
    Returns (expected photons (Nw,Nh) WITHOUT background/QE, true_data dict).
    """
    exposure_time = exposure_time or params["exposure_time"]
    psf = psf or PsfTables(params)
    Nw, Nh = params["image_size"]
    pl = params["pixel_length"] / params["magnification"]
    p_0 = numpy.asarray(params["focal_point"], dtype=float)
    times = [t for t, _ in input_data]
    windows, exposure_time = frame_windows(times, frame_index, start_time, exposure_time, params)
    expected = numpy.zeros((Nw, Nh))
    optinfo = {}
    if fluorescence_states is not None and params["bleaching_switch"]:
        beta, N_emit0 = photon_budget_scale(params)
    for k, unit_time in windows:
        particles = input_data[k][1]
        plane = numpy.zeros((Nw, Nh))
        for row in particles:
            x, y, z, m_id, p_state = row
            m_id = int(m_id)
            p_i = numpy.array([x, y, z])
            depth = (p_i - p_0)[0]
            N_emit = emitted(params, depth, unit_time)
            if fluorescence_states is not None and params["bleaching_switch"]:
                if m_id not in fluorescence_states:
                    fluorescence_states[m_id] = budget_draw(m_id) * beta * N_emit0
                budget = fluorescence_states[m_id] - N_emit
                if budget <= 0:
                    budget = 0
                    p_state = 0.0
                fluorescence_states[m_id] = budget
            normalization = spot_weight(params, N_emit, p_state)
            if normalization > 0.0:
                rel = p_i - p_0
                if exact:
                    _, table = psf.get(rel[0])
                    overlay_signal_exact(plane, table, rel, pl, normalization)
                else:
                    _, sat = psf.sat(rel[0])
                    overlay_signal_sat(plane, sat, rel, pl, normalization)
            if m_id not in optinfo:
                optinfo[m_id] = numpy.zeros(8, dtype=numpy.float64)
            optinfo[m_id] += numpy.array([
                unit_time, unit_time * p_state, unit_time * p_i[1], unit_time * p_i[2],
                unit_time * p_i[1], unit_time * p_i[2], unit_time * depth, normalization])
        expected += plane
    for m_id in optinfo:
        optinfo[m_id][1] /= exposure_time
        optinfo[m_id][2: 6] /= optinfo[m_id][0]
        optinfo[m_id][2] = (optinfo[m_id][2] - p_0[1]) / pl + (Nw - 1) * 0.5
        optinfo[m_id][3] = (optinfo[m_id][3] - p_0[2]) / pl + (Nh - 1) * 0.5
    return expected, optinfo


# ------------------------------------------------------------------ detector (a17-a22)
def detector_expectation(photons, params):
    """``_epifm.py:1434-1441``: expected photoelectrons = QE * (photons + background)."""
    photons = numpy.array(photons, dtype=float)
    if params["background_switch"]:
        photons = photons + params["background_mean"]
    return params["QE"] * photons


def emccd_probability(S, E, a):
    """``_epifm.py:365-369``."""
    from scipy.special import i1e
    X = a * S
    Y = 2 * numpy.sqrt(E * X)
    return 1.0 / Y * numpy.exp(-X + Y) * i1e(Y)


def emccd_pmf(expected, emgain):
    """``_epifm.py:372-391``: support S and normalised pmf for one pixel."""
    sigma = numpy.sqrt(expected) * 5 + 10
    s_min = max(0, emgain * int(expected - sigma))
    s_max = emgain * int(expected + sigma)
    S = numpy.arange(s_min, s_max)
    a = 1.0 / emgain
    if S[0] > 0:
        p = emccd_probability(S, expected, a)
    else:
        p = numpy.zeros(len(S))
        p[0] = 1.0 / (2 * a * expected)
        p[1:] = emccd_probability(S[1:], expected, a)
    p /= p.sum()
    return S, p


def cmos_readout_pmf(table):
    """``_epifm.py:334-339``: (values, normalised probabilities) from the RNDist table."""
    table = numpy.asarray(table, dtype=float)
    return table[:, 0], table[:, 1] / table[:, 1].sum()


def detector_draw(expected, params, rng, cmos_table=None):
    """Stochastic stage with a numpy RandomState, same draw ORDER as the reference
    (noise array first, then per-pixel signal; ``_epifm.py:1445-1457``)."""
    kind = params["detector_type"]
    shape = expected.shape
    if kind == "CMOS":
        vals, p = cmos_readout_pmf(cmos_table)
        noise = rng.choice(vals, size=shape, p=p)
        signal = rng.poisson(expected).astype(float)
    elif kind in ("EMCCD", "CCD"):
        rn = params["readout_noise"]
        noise = rng.normal(0, rn, shape) if rn > 0 else numpy.zeros(shape)
        if kind == "CCD":
            signal = rng.poisson(expected).astype(float)
        else:
            flat = expected.ravel()
            signal = numpy.zeros(flat.size)
            for i in range(flat.size):
                if flat[i] > 0:
                    S, p = emccd_pmf(flat[i], params["emgain"])
                    signal[i] = rng.choice(S, None, p=p)
            signal = signal.reshape(shape)
    else:
        raise RuntimeError("Unknown detector type was given [{}].".format(kind))
    return signal, noise


def adc_params(params, column_or_pixel_normals=None):
    """``_epifm.py:926-963``: (offset, gain) maps (Nw,Nh).  For 'pixel' / 'column' FPN
    the caller passes the N(ADC0, count) draws (Nw*Nh or Nh values)."""
    Nw, Nh = params["image_size"]
    ADC0 = params["adc_offset"]
    kind = params["fpn_type"]
    if kind == 'none':
        offset = numpy.full(Nw * Nh, ADC0)
    elif kind == 'pixel':
        offset = numpy.rint(numpy.asarray(column_or_pixel_normals, dtype=float).reshape(Nw * Nh))
    elif kind == 'column':
        column = numpy.asarray(column_or_pixel_normals, dtype=float).reshape(Nh)
        offset = numpy.rint(numpy.tile(column, (1, Nw)).reshape(Nh * Nw))
    else:
        raise ValueError("FPN type [{}] is invalid ['pixel', 'column' or 'none']".format(kind))
    gain = (params["adc_fullwell"] - 0.0) / (pow(2.0, params["adc_bit"]) - offset)
    return offset.reshape([Nw, Nh]), gain.reshape([Nw, Nh])


def adc_counts(photoelectron, fullwell, gain, offset, bit):
    """``_epifm.py:1472-1484``: no rounding, fractional counts."""
    ADC = numpy.array(photoelectron, dtype=float)
    ADC[ADC > fullwell] = fullwell
    ADC_max = 2 ** bit - 1
    ADC /= gain
    ADC += offset
    ADC[ADC > ADC_max] = ADC_max
    ADC[ADC < 0] = 0
    return ADC


# ------------------------------------------------------------------ sampling (a1-a4)
def move_points(rng, points, D, dt, ndim=3):
    """``sampling.py:88-119`` with the per-coordinate draws vectorised in the same
    (particle-major, axis-minor) order, so a seeded RandomState gives identical output."""
    if D is None:
        D = numpy.zeros(ndim)
    elif not numpy.iterable(D):
        D = numpy.ones(ndim) * D
    else:
        D = numpy.asarray(D)
    ret = points.copy()
    z = rng.normal(0.0, 1.0, size=(len(ret), ndim))
    # rng.normal(0, s) == 0 + s * standard_normal (numpy legacy distributions)
    ret[:, :ndim] += z * numpy.sqrt(2 * D * dt)[None, :]
    return ret


# ------------------------------------------------------------ whole frame with a RandomState
def output_frame(input_data, params, rng, frame_index=0, start_time=0.0, exposure_time=None,
                 fluorescence_states=None, cmos_table=None, adc_normals=None, psf=None, exact=False):
    """``_EPIFMSimulator.output_frame`` (``_epifm.py:1121-1225``) driven by a numpy
    RandomState in the reference's draw order (photon budgets in particle order ->
    readout noise array -> per-pixel signal), so that with the same seed the result equals
    the reference's bit for bit (CCD / CMOS / EMCCD).  Returns (camera (Nw,Nh,2), true_data)."""
    def budget_draw(_m_id):
        return rng.exponential(scale=1.0)

    photons, true_data = expected_frame(
        input_data, params, frame_index=frame_index, start_time=start_time, exposure_time=exposure_time,
        fluorescence_states=fluorescence_states, budget_draw=budget_draw, psf=psf, exact=exact)
    expected = detector_expectation(photons, params)
    signal, noise = detector_draw(expected, params, rng, cmos_table=cmos_table)
    offset, gain = adc_params(params, adc_normals)
    camera = numpy.zeros(expected.shape + (2,))
    camera[:, :, 0] = expected
    camera[:, :, 1] = adc_counts(signal + noise, params["adc_fullwell"], gain, offset, params["adc_bit"])
    return camera, true_data
