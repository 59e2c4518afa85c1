"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import the UNMODIFIED reference (``/root/reference/src/scopyon``) in the build
container, where two of its off-path dependencies (``pint``, ``hmmlearn``) are
not installed.  We register tiny stand-in modules for those two names before
the import; nothing under ``/root/reference`` is copied or modified.

The stand-ins cover exactly the surface the reference touches:
  * ``scopyon/constants.py:1-17``   -> ``pint.UnitRegistry`` with ``Quantity``,
    ``avogadro_number``, ``planck_constant``, ``speed_of_light``, ``joule``,
    ``meter``;
  * ``scopyon/config.py:10-11,108-114,133-145`` -> ``Quantity(value, units)``,
    ``.to_base_units()``, ``.units``, ``.magnitude``, ``.check()``,
    ``pint.Quantity``, ``pint.errors.DimensionalityError``;
  * ``scopyon/analysis/hmm.py:3-4`` -> ``hmmlearn.base.BaseHMM``, ``hmmlearn._utils``.

``/root/reference`` does not exist on the GPU box: callers must use
``reference_available()`` and skip when it returns False.
"""
import os
import re
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "scopyon"))


# ---------------------------------------------------------------------------
# a minimal unit algebra: a unit is (scale to SI, dimension exponents dict)
_DIM_BASE = {
    "m": (1.0, {"L": 1}), "meter": (1.0, {"L": 1}), "nm": (1e-9, {"L": 1}),
    "um": (1e-6, {"L": 1}), "mm": (1e-3, {"L": 1}), "cm": (1e-2, {"L": 1}),
    "s": (1.0, {"T": 1}), "second": (1.0, {"T": 1}), "ms": (1e-3, {"T": 1}),
    "us": (1e-6, {"T": 1}), "kg": (1.0, {"M": 1}), "g": (1e-3, {"M": 1}),
    "radian": (1.0, {}), "rad": (1.0, {}), "degree": (3.141592653589793 / 180.0, {}),
    "dimensionless": (1.0, {}), "": (1.0, {}),
    "joule": (1.0, {"M": 1, "L": 2, "T": -2}), "J": (1.0, {"M": 1, "L": 2, "T": -2}),
    "W": (1.0, {"M": 1, "L": 2, "T": -3}),
}


class _Unit:
    def __init__(self, scale=1.0, dims=None, text="dimensionless"):
        self.scale = scale
        self.dims = {k: v for k, v in (dims or {}).items() if v != 0}
        self.text = text

    @staticmethod
    def parse(text):
        if isinstance(text, _Unit):
            return text
        src = str(text).strip()
        flat = re.sub(r"\s*\^\s*", "^", src.replace("**", "^"))
        toks = flat.replace("*", " * ").replace("/", " / ").split()
        scale, dims, sign = 1.0, {}, 1
        for tok in toks:
            if tok == "*":
                sign = 1
                continue
            if tok == "/":
                sign = -1
                continue
            name, _, power = tok.partition("^")
            power = float(power) if power else 1.0
            if name not in _DIM_BASE:
                raise ValueError("unknown unit [{}]".format(name))
            s, d = _DIM_BASE[name]
            scale *= s ** (sign * power)
            for k, v in d.items():
                dims[k] = dims.get(k, 0) + sign * power * v
        return _Unit(scale, dims, src or "dimensionless")

    def base(self):
        names = {"L": "m", "M": "kg", "T": "s"}
        num = [names[k] if v == 1 else "{} ** {:g}".format(names[k], v)
               for k, v in sorted(self.dims.items()) if v > 0]
        den = [names[k] if v == -1 else "{} ** {:g}".format(names[k], -v)
               for k, v in sorted(self.dims.items()) if v < 0]
        text = " * ".join(num) if num else ("1" if den else "dimensionless")
        if den:
            text += " / " + " / ".join(den)
        if not self.dims:
            text = self.text if self.scale == 1.0 else "radian" if "deg" in self.text else "dimensionless"
        return _Unit(1.0, self.dims, text)

    def __eq__(self, other):
        other = _Unit.parse(other)
        return self.dims == other.dims and self.scale == other.scale and (
            bool(self.dims) or self.text == other.text)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __mul__(self, other):
        other = _Unit.parse(other)
        d = dict(self.dims)
        for k, v in other.dims.items():
            d[k] = d.get(k, 0) + v
        return _Unit(self.scale * other.scale, d, "{} * {}".format(self.text, other.text))

    def __format__(self, spec):
        return self.text

    def __repr__(self):
        return self.text


class DimensionalityError(TypeError):
    def __init__(self, units1, units2, dim1=None, dim2=None):
        super().__init__("Cannot convert from '{}' to '{}'".format(units1, units2))


class _Quantity:
    def __init__(self, value, units="dimensionless"):
        if isinstance(value, _Quantity):
            value, units = value.magnitude, value.units
        self.magnitude = value
        self.units = _Unit.parse(units)

    @property
    def dimensionality(self):
        return self.units.dims

    def _scaled(self, factor):
        import numpy
        if isinstance(self.magnitude, (list, tuple)):
            return (numpy.asarray(self.magnitude, dtype=float) * factor).tolist() if factor != 1.0 \
                else self.magnitude
        return self.magnitude * factor if factor != 1.0 else self.magnitude

    def to_base_units(self):
        return _Quantity(self._scaled(self.units.scale), self.units.base())

    def to(self, units):
        units = _Unit.parse(units)
        if units.dims != self.units.dims:
            raise DimensionalityError(self.units, units)
        return _Quantity(self._scaled(self.units.scale / units.scale), units)

    def check(self, other):
        other = other.units if isinstance(other, _Quantity) else _Unit.parse(other)
        return self.units.dims == other.dims

    def __mul__(self, other):
        if isinstance(other, _Quantity):
            return _Quantity(self.magnitude * other.magnitude, self.units * other.units)
        if isinstance(other, _Unit):
            return _Quantity(self.magnitude, self.units * other)
        return _Quantity(self.magnitude * other, self.units)

    __rmul__ = __mul__


class _UnitRegistry:
    Quantity = _Quantity

    def __init__(self):
        self.avogadro_number = _Quantity(6.02214076e23, "dimensionless")
        self.planck_constant = _Quantity(6.62607015e-34, "joule * s")
        self.speed_of_light = _Quantity(299792458.0, "m / s")
        self.joule = _Unit.parse("joule")
        self.meter = _Unit.parse("m")


def _install_stand_ins():
    if "pint" not in sys.modules:
        try:
            import pint  # noqa: F401  (real pint wins when present)
        except ImportError:
            pint = types.ModuleType("pint")
            pint.UnitRegistry = _UnitRegistry
            pint.Quantity = _Quantity
            errors = types.ModuleType("pint.errors")
            errors.DimensionalityError = DimensionalityError
            pint.errors = errors
            pint.__stand_in__ = True
            sys.modules["pint"] = pint
            sys.modules["pint.errors"] = errors
    if "hmmlearn" not in sys.modules:
        try:
            import hmmlearn  # noqa: F401
        except ImportError:
            hmmlearn = types.ModuleType("hmmlearn")
            base = types.ModuleType("hmmlearn.base")
            base.BaseHMM = type("BaseHMM", (), {})
            utils = types.ModuleType("hmmlearn._utils")
            hmmlearn.base, hmmlearn._utils = base, utils
            hmmlearn.__stand_in__ = True
            sys.modules["hmmlearn"] = hmmlearn
            sys.modules["hmmlearn.base"] = base
            sys.modules["hmmlearn._utils"] = utils


def import_reference():
    """Return the live reference package ``scopyon`` (unmodified)."""
    if not reference_available():
        raise ImportError("reference tree {} is not present on this machine".format(REFERENCE_SRC))
    _install_stand_ins()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import scopyon
    return scopyon
