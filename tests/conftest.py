import json
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree /root/reference")


def pytest_collection_modifyitems(config, items):
    import ref_shim
    have_ref = ref_shim.reference_available()
    skip_ref = pytest.mark.skip(reason="/root/reference is not present on this machine")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


def golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".json"):
        with open(path) as f:
            return json.load(f)
    return numpy.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def known_answers():
    return golden("known_answers.json")


def make_configs(yaml_update=None, seed=0):
    """(Configuration, EPIFMConfigs, oracle parameter dict) of the product's host layer."""
    import warnings
    import scopyon_b200
    from scopyon_b200 import _epifm
    config = scopyon_b200.DefaultConfiguration()
    if yaml_update:
        config.update(yaml_update)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        configs = _epifm.EPIFMConfigs(config.default, rng=numpy.random.RandomState(seed))
    return config, configs, configs.as_oracle_params()


def format_inputs(config, inputs):
    import warnings
    import scopyon_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sim = scopyon_b200.EPIFMSimulator(config=config, method="default", rng=numpy.random.RandomState(0))
    return sim._EPIFMSimulator__format_inputs(inputs)


def cmos_table():
    from scopyon_b200._epifm import catalog_tables
    rn = catalog_tables()["cmos_readout"]
    return numpy.stack([rn["electrons"], rn["weight"]], axis=1)


def gpu_engine(yaml_update=None, precision="f64", seed=0):
    """(config, configs, oracle params, DeviceEngine) on cuda:0."""
    from scopyon_b200.engine import DeviceEngine, SatStore
    config, configs, params = make_configs(yaml_update, seed=seed)
    SatStore.clear_shared()     # every test starts from an empty table store
    return config, configs, params, DeviceEngine(configs, precision=precision)


def oracle_tables(params, engine, keys):
    """Build the engine's tables for ``keys`` from the ORACLE's radial profiles (so that
    device and C oracle start from bit-identical input) and return the C-oracle twins:
    (sats (n, pitch, pitch) int64, inv_scale (n,), slot_of_key)."""
    import torch
    import c_oracle
    import epifm_oracle as orc
    keys = [int(k) for k in keys]
    profs = numpy.stack([orc.radial_profile(params, engine.table_depth(k)) for k in keys])
    first = engine._build_tables(keys, radial=torch.from_numpy(profs).to(engine.device))
    for i, k in enumerate(keys):
        engine.slot_host[k] = first + i
    engine.slot_of_key.copy_(torch.from_numpy(engine.slot_host))
    sats, inv = [], []
    for prof in profs:
        S, s = c_oracle.sat_from_table(c_oracle.table_from_radial(prof))
        sats.append(S)
        inv.append(s)
    return numpy.stack(sats), numpy.array(inv), engine.slot_host.copy()
