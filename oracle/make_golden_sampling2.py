"""TEST INFRASTRUCTURE ONLY.  Golden vectors of the reference's multi-state trajectory generator
(``/root/reference/src/scopyon/sampling2.py``) for machines without the reference tree:

    python oracle/make_golden_sampling2.py      ->  tests/golden/sampling2_case.npz

The reference runs unmodified; its transition step (``searchsorted(..., side='leff')``, ``:66``, a
ValueError as written) is executed with that keyword read as ``'left'`` by wrapping
``numpy.searchsorted`` for the duration of the call."""
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref_shim.import_reference()
    from scopyon import sampling2 as R
    real = numpy.searchsorted
    numpy.searchsorted = lambda arr, v, side='left', sorter=None: real(
        arr, v, side='left' if side == 'leff' else side, sorter=sorter)
    try:
        lower, upper = numpy.array([0.0, -1e-6, 2e-7]), numpy.array([1e-6, 1e-6, 2e-7])
        D = numpy.array([1e-12, 0.0, 3e-13])
        transmat = numpy.array([[0.0, 2.0, 0.5], [1.0, 0.0, 0.0], [0.3, 4.0, 0.0]])
        t = numpy.array([0.0, 0.1, 0.1, 0.25, 0.5])
        out = R.sample(t, [40, 25, 10], lower=lower, upper=upper, D=D, transmat=transmat, ndim=3, periodic=True,
                       rng=numpy.random.RandomState(2024))
        free = R.sample(t, [30, 30], lower=0.0, upper=1e-6, D=[2e-13, 5e-12], ndim=2, periodic=False,
                        rng=numpy.random.RandomState(7))
    finally:
        numpy.searchsorted = real
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "sampling2_case.npz")
    numpy.savez_compressed(path, t=t, lower=lower, upper=upper, D=D, transmat=transmat,
                           periodic_switching=numpy.stack(out), free=numpy.stack(free))
    print("wrote", path, numpy.stack(out).shape, numpy.stack(free).shape)


if __name__ == "__main__":
    main()
