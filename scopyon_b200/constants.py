"""Physical constants used on the path (reference ``constants.py:6-17``).

The reference derives them from pint's registry (CODATA 2018, exact SI values);
the numbers are spelled out here so that nothing depends on pint.
"""
from .units import Quantity, Unit

__all__ = ['ureg', 'Quantity', 'Q_', 'AVOGADROS_NUMBER', 'N_A', 'hc']

Q_ = Quantity


class _Registry:
    """The few attributes user scripts take from ``scopyon.constants.ureg``."""
    Quantity = Quantity

    def __getattr__(self, name):
        return Unit.parse(name)


ureg = _Registry()

# Avogadro constant [1/mol]
AVOGADROS_NUMBER = 6.02214076e+23
N_A = AVOGADROS_NUMBER

# (Planck constant) * (speed of light) [J m]
hc = 6.62607015e-34 * 299792458
