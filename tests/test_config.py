"""Configuration boundary: the reference's own config tests (test/test_config.py:12-82),
run against scopyon_b200's pint-free implementation."""
import warnings

import pytest

from scopyon_b200 import Configuration, DefaultConfiguration
from scopyon_b200.constants import Q_
from scopyon_b200.units import DimensionalityError


def test_default_loads_and_exposure_roundtrip():
    config = DefaultConfiguration()
    assert config.default.detector.exposure_time == pytest.approx(0.1)
    config.default.detector.exposure_time = 0.033
    assert config.default.detector.exposure_time == pytest.approx(0.033)


def test_update_without_and_with_units():
    config = DefaultConfiguration()
    config.update("""
    default:
        detector:
            exposure_time: 0.033
    """)
    assert config.default.detector.exposure_time == pytest.approx(0.033)
    config.update("""
    default:
        detector:
            exposure_time:
                value: 0.1
    """)
    assert config.default.detector.exposure_time == pytest.approx(0.1)
    config.update("""
    default:
        detector:
            exposure_time:
                value: 100
                units: ms
    """)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        assert config.default.detector.exposure_time == pytest.approx(0.1)
    assert any("Unit conversion" in str(w.message) for w in caught)


def test_quantity_assignment_and_dimension_check():
    config = DefaultConfiguration()
    config.default.detector.exposure_time = Q_(33, 'ms')
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert config.default.detector.exposure_time == pytest.approx(0.033)
    with pytest.raises(DimensionalityError):
        config.default.detector.exposure_time = Q_(33, 'm')


def test_mapping_protocol_and_errors(tmp_path):
    config = DefaultConfiguration()
    detector = dict(**config.default.detector)       # leaves only, SI magnitudes
    assert detector["pixel_length"] == 16.0e-6 and detector["image_size"] == [512, 512]
    assert "fluorophore" not in list(config.default)  # sub-trees are not leaves
    with pytest.raises(KeyError):
        config.default.no_such_key
    with pytest.raises(KeyError):
        config.default.detector.no_such_key = 1
    with pytest.raises(TypeError):
        config.default.detector.exposure_time = dict(value=1)
    with pytest.raises(ValueError):
        config.default.detector = 3
    path = tmp_path / "saved.yaml"
    config.default.magnification = 360
    config.save(path)
    again = Configuration(filename=str(path))
    assert again.default.magnification == 360
    assert again.default.light_source.flux_density == 400000.0


def test_reference_yaml_tags_load():
    # user files written for scopyon carry explicit !!bool / !!int tags (scopyon.yaml:3,70)
    config = DefaultConfiguration()
    config.update("environ: {processes: !!int 4}\ndefault: {dichroic_mirror: {switch: !!bool false}}")
    assert config.environ.processes == 4 and config.default.dichroic_mirror.switch is False
