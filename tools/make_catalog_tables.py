"""Regenerate ``scopyon_b200/data/catalog_tables.json`` from the reference catalog.

The image-formation path reads exactly two things from the reference's 1 330-file
catalog:
  * per fluorophore: the index of the emission peak on the integer-nm wavelength grid
    (``_epifm.py:839-858`` -> ``psf_wavelength``) and ``sum(fluoem_norm)``
    (``_epifm.py:1309,1314``);
  * the CMOS read-noise distribution ``catalog/detector/RNDist_F40.csv``
    (``_epifm.py:334-339``).
This script derives those numbers by running the unmodified reference here; the
product ships the derived table, not the catalog.  Run: ``python tools/make_catalog_tables.py``.
"""
import json
import os
import sys
import warnings

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402


def main():
    scopyon = ref_shim.import_reference()
    from scopyon import _epifm, io
    catalog_dir = os.path.join(os.path.dirname(scopyon.__file__), "catalog")
    names = sorted(os.path.splitext(f)[0] for f in os.listdir(os.path.join(catalog_dir, "fluorophore"))
                   if f.endswith(".csv"))
    grid_min, grid_max = 300.0e-9, 1000.0e-9  # scopyon.yaml:43-48
    grid = numpy.arange(grid_min, grid_max, 1e-9, dtype=float)
    fluor = {}
    for name in names:
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ex, em = io.read_fluorophore_catalog(name)
                em = numpy.array(_epifm.EPIFMConfigs.calculate_efficiency(em, grid))
        except Exception as exc:  # a few catalog files are malformed in the reference too
            print("skip", name, type(exc).__name__, exc)
            continue
        index_em = int(em.argmax())
        em[index_em] = 100
        em /= sum(em)
        fluor[name] = dict(index_em=index_em, fluoem_norm_sum=float(numpy.sum(em)))
    rn = numpy.loadtxt(os.path.join(catalog_dir, "detector", "RNDist_F40.csv"), delimiter=",")
    out = dict(
        wavelength_grid=dict(min=grid_min, max=grid_max, step=1e-9),
        fluorophore=fluor,
        cmos_readout=dict(electrons=rn[:, 0].tolist(), weight=rn[:, 1].tolist()),
    )
    path = os.path.join(ROOT, "scopyon_b200", "data", "catalog_tables.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, len(fluor), "fluorophores,", rn.shape[0], "read-noise rows")


if __name__ == "__main__":
    main()
