"""Unit handling for the YAML configuration boundary.

The reference converts ``{value, units}`` leaves to SI magnitudes with ``pint``
(``/root/reference/src/scopyon/config.py:101-117``, ``constants.py:6-9``).  ``pint``
is not part of this image, and the path needs only a handful of SI-coherent
units, so this module is a small self-contained quantity type with the same
surface the configuration layer uses: ``Quantity(value, units)``, ``.magnitude``,
``.units``, ``.to_base_units()``, ``.to(units)``, ``.check(other)`` and
``DimensionalityError``.
"""
import math
import re

__all__ = ["Unit", "Quantity", "Q_", "DimensionalityError"]

# name -> (factor to SI base, {base dimension: exponent}); dimensions are (m, kg, s)
_PREFIX = {"": 1.0, "k": 1e3, "c": 1e-2, "m": 1e-3, "u": 1e-6, "µ": 1e-6, "n": 1e-9, "p": 1e-12}
_ROOTS = {
    "m": (1.0, (1, 0, 0)), "meter": (1.0, (1, 0, 0)), "metre": (1.0, (1, 0, 0)),
    "g": (1e-3, (0, 1, 0)), "gram": (1e-3, (0, 1, 0)),
    "s": (1.0, (0, 0, 1)), "sec": (1.0, (0, 0, 1)), "second": (1.0, (0, 0, 1)),
    "J": (1.0, (2, 1, -2)), "joule": (1.0, (2, 1, -2)),
    "W": (1.0, (2, 1, -3)), "watt": (1.0, (2, 1, -3)),
    "rad": (1.0, (0, 0, 0)), "radian": (1.0, (0, 0, 0)),
    "deg": (math.pi / 180.0, (0, 0, 0)), "degree": (math.pi / 180.0, (0, 0, 0)),
    "dimensionless": (1.0, (0, 0, 0)),
}
_BASE_NAMES = ("m", "kg", "s")


class DimensionalityError(TypeError):
    """Raised when a quantity of the wrong dimension is assigned or converted."""

    def __init__(self, units1, units2, dim1=None, dim2=None):
        self.units1, self.units2 = units1, units2
        super().__init__("Cannot convert from '{}' to '{}'".format(units1, units2))


def _lookup(name):
    if name in _ROOTS:
        return _ROOTS[name]
    if name == "kg":
        return 1.0, (0, 1, 0)
    for prefix, factor in _PREFIX.items():
        root = name[len(prefix):]
        if prefix and name.startswith(prefix) and root in _ROOTS and root not in ("dimensionless",):
            scale, dims = _ROOTS[root]
            return scale * factor, dims
    raise ValueError("unknown unit '{}'".format(name))


class Unit:
    """A product of powers of SI base units with a scale factor."""

    def __init__(self, factor=1.0, dims=(0, 0, 0), text="dimensionless"):
        self.factor = float(factor)
        self.dims = tuple(dims)
        self.text = text

    @classmethod
    def parse(cls, spec):
        if isinstance(spec, Unit):
            return spec
        if spec is None:
            return cls()
        text = str(spec).strip()
        flat = re.sub(r"\s*\^\s*", "^", text.replace("**", "^"))
        factor, dims, sign = 1.0, [0.0, 0.0, 0.0], 1.0
        for tok in re.findall(r"[*/]|[^\s*/]+", flat):
            if tok == "*":
                sign = 1.0
            elif tok == "/":
                sign = -1.0
            else:
                name, _, power = tok.partition("^")
                if name in ("1", ""):
                    continue
                power = sign * (float(power) if power else 1.0)
                scale, d = _lookup(name)
                factor *= scale ** power
                dims = [a + power * b for a, b in zip(dims, d)]
        return cls(factor, dims, text or "dimensionless")

    @property
    def dimensionless(self):
        return not any(self.dims)

    def to_base(self):
        if self.dimensionless:
            # pint keeps 'radian' as a base unit and reduces degrees to it
            angular = re.search(r"rad|deg", self.text) is not None
            return Unit(1.0, self.dims, "radian" if angular else "dimensionless")
        num, den = [], []
        for name, power in zip(_BASE_NAMES, self.dims):
            if power > 0:
                num.append(name if power == 1 else "{} ** {:g}".format(name, power))
            elif power < 0:
                den.append(name if power == -1 else "{} ** {:g}".format(name, -power))
        text = " * ".join(num) if num else "1"
        if den:
            text += " / " + " / ".join(den)
        return Unit(1.0, self.dims, text)

    def same_dimension(self, other):
        return all(abs(a - b) < 1e-12 for a, b in zip(self.dims, Unit.parse(other).dims))

    def __eq__(self, other):
        try:
            other = Unit.parse(other)
        except ValueError:
            return False
        if not self.same_dimension(other) or not math.isclose(self.factor, other.factor, rel_tol=1e-15):
            return False
        return (not self.dimensionless) or self.to_base().text == other.to_base().text \
            and math.isclose(self.factor, other.factor)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.dims, self.factor))

    def __mul__(self, other):
        other = Unit.parse(other)
        return Unit(self.factor * other.factor, [a + b for a, b in zip(self.dims, other.dims)],
                    "{} * {}".format(self.text, other.text))

    def __format__(self, spec):
        return self.text

    def __str__(self):
        return self.text

    __repr__ = __str__


def _scale(value, factor):
    if factor == 1.0 or value is None:
        return value
    if isinstance(value, (list, tuple)):
        return [v * factor for v in value]
    return value * factor


class Quantity:
    """A magnitude (scalar, list or ndarray) with a ``Unit``."""

    def __init__(self, value, units=None):
        if isinstance(value, Quantity):
            units = value.units if units is None else units
            value = value.magnitude
        self.magnitude = value
        self.units = Unit.parse(units)

    @property
    def dimensionality(self):
        return dict(zip(_BASE_NAMES, self.units.dims))

    def to_base_units(self):
        return Quantity(_scale(self.magnitude, self.units.factor), self.units.to_base())

    def to(self, units):
        units = Unit.parse(units)
        if not self.units.same_dimension(units):
            raise DimensionalityError(self.units, units)
        return Quantity(_scale(self.magnitude, self.units.factor / units.factor), units)

    def check(self, other):
        other = other.units if isinstance(other, Quantity) else other
        return self.units.same_dimension(other)

    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.magnitude * other.magnitude, self.units * other.units)
        if isinstance(other, Unit):
            return Quantity(self.magnitude, self.units * other)
        return Quantity(self.magnitude * other, self.units)

    __rmul__ = __mul__

    def __repr__(self):
        return "<Quantity({}, '{}')>".format(self.magnitude, self.units)


Q_ = Quantity
