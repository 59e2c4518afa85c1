"""YAML configuration tree with unit-aware leaves.

Mirrors the interface of the reference's ``scopyon.config``
(``/root/reference/src/scopyon/config.py:26-165``): ``Configuration`` is a read-only
``Mapping`` over a nested dict; attribute access descends into sub-trees and
converts ``{value, units}`` leaves to SI magnitudes; attribute assignment replaces
a leaf (checking the dimension when a ``Quantity`` is given); ``update()`` deep-merges
another configuration, a dict or a YAML string.  This layer is host-side glue
(microseconds per call) and is not accelerated.
"""
import collections.abc
import os
import pathlib
import warnings

import yaml as _yaml

from .units import DimensionalityError, Quantity

__all__ = ["Configuration", "DefaultConfiguration"]

_Loader = getattr(_yaml, "CSafeLoader", _yaml.SafeLoader)
_DEFAULT_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "default_config.yaml")


def _parse(text):
    # the reference's files carry explicit ``!!bool`` / ``!!int`` tags, which the safe
    # loader understands, so user YAML written for scopyon loads as is.
    return _yaml.load(text, Loader=_Loader)


def _merge(dst, src):
    """Deep-merge ``src`` into ``dst`` (reference ``dict_merge``, ``config.py:19-24``)."""
    for key, val in src.items():
        if isinstance(dst.get(key), dict) and isinstance(val, collections.abc.Mapping):
            _merge(dst[key], val)
        else:
            dst[key] = val


def _is_leaf(node):
    return not isinstance(node, dict) or 'value' in node


class Configuration(collections.abc.Mapping):

    def __init__(self, filename=None, yaml=None):
        tree = None
        if filename is not None:
            assert yaml is None
            with open(filename) as f:
                tree = _parse(f.read())
        elif yaml is not None:
            tree = yaml
        object.__setattr__(self, '_tree', tree)

    # -- serialisation
    def __repr__(self):
        return _yaml.dump(self._tree, default_flow_style=False)

    def save(self, file):
        """Save the configuration as a YAML file."""
        assert isinstance(file, (str, pathlib.PurePath))
        with open(str(file), 'w') as f:
            f.write(repr(self))

    def update(self, conf):
        if isinstance(conf, Configuration):
            _merge(self._tree, conf.yaml)
        elif isinstance(conf, dict):
            _merge(self._tree, conf)
        else:
            _merge(self._tree, _parse(conf))

    @property
    def yaml(self):
        return self._tree

    # -- Mapping protocol: iteration yields leaves only, so ``**config.detector`` expands
    #    to keyword arguments of SI magnitudes (``_epifm.py:778-790``)
    def get(self, key, defaultobj=None):
        return self._tree.get(key, defaultobj)

    def __getitem__(self, key):
        return getattr(self, key)

    def __len__(self):
        return len(self._tree)

    def __iter__(self):
        return (key for key, node in self._tree.items() if _is_leaf(node))

    def __getattr__(self, key):
        tree = object.__getattribute__(self, '_tree')
        if tree is None or key not in tree:
            raise KeyError("'{}'".format(key))
        node = tree[key]
        if not isinstance(node, dict):
            return node
        if 'value' not in node:
            return Configuration(yaml=node)
        if node['value'] is not None and 'units' in node:
            given = Quantity(node['value'], node['units'])
            si = given.to_base_units()
            if given.units != si.units:
                warnings.warn("Unit conversion in '{}' from [{}] to [{}]".format(key, given.units, si.units))
            return si.magnitude
        return node['value']

    def __setattr__(self, key, value):
        if key.startswith('_'):
            object.__setattr__(self, key, value)
            return
        if key not in self._tree:
            raise KeyError("'{}'".format(key))
        if isinstance(value, dict):
            raise TypeError("The given value for '{}' has wrong type: {}".format(key, value))
        node = self._tree[key]
        if isinstance(node, dict) and 'value' not in node:
            raise ValueError("Cannot update '{}'.".format(key))
        if isinstance(value, Quantity):
            if isinstance(node, dict) and 'units' in node:
                if not value.check(Quantity(node['value'], node['units'])):
                    raise DimensionalityError(value.units, node['units'])
            self._tree[key] = dict(value=value.magnitude, units=str(value.units))
        else:
            self._tree[key] = value  # a bare number is taken as SI; 'units' is dropped


class DefaultConfiguration(Configuration):
    """The packaged defaults (``default_config.yaml``)."""

    def __init__(self):
        with open(_DEFAULT_FILE) as f:
            Configuration.__init__(self, yaml=_parse(f.read()))
