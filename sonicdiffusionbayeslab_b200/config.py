"""OmegaConf-subset configuration objects (omegaconf is not installable here).

Covers what the reference uses (/root/reference/main.py:11-14 and the experiment drivers):
``load(path)``, attribute access, ``.get(key, default)``, item access, ``to_container``.  Missing
attributes raise ``AttributeError`` like OmegaConf >= 2.1; ``DEFAULTS`` supplies the keys the
reference code reads but its shipped YAML files omit.
"""
from __future__ import annotations

import yaml

# keys read by attribute in the reference drivers but absent from configs/*.yaml
DEFAULTS = {
    "dpm_solver": {"experiment_params": {"algorithm_type": "dpmsolver++", "final_sigmas_type": "zero"}},
    "skip_steps": {"experiment_params": {"solver_order": 2, "algorithm_type": "dpmsolver++",
                                         "final_sigmas_type": "zero"}},
}


class DictConfig:
    def __init__(self, data: dict):
        object.__setattr__(self, "_data", {k: _wrap(v) for k, v in data.items()})

    def __getattr__(self, key):
        try:
            return self._data[key]
        except KeyError as e:
            raise AttributeError(f"Missing key {key}") from e

    def __setattr__(self, key, value):
        self._data[key] = _wrap(value)

    def __getitem__(self, key):
        return self._data[key]

    def __contains__(self, key):
        return key in self._data

    def __iter__(self):
        return iter(self._data)

    def keys(self):
        return self._data.keys()

    def items(self):
        return self._data.items()

    def get(self, key, default=None):
        v = self._data.get(key, default)
        return default if v is None else v

    def __repr__(self):
        return f"DictConfig({to_container(self)!r})"


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, dict):
        return DictConfig(v)
    if isinstance(v, list) and not isinstance(v, ListConfig):
        return ListConfig(_wrap(x) for x in v)
    return v


def to_container(cfg, resolve=True):
    if isinstance(cfg, DictConfig):
        return {k: to_container(v) for k, v in cfg._data.items()}
    if isinstance(cfg, list):
        return [to_container(v) for v in cfg]
    return cfg


def _merge_defaults(data: dict, defaults: dict):
    for k, v in defaults.items():
        if isinstance(v, dict):
            _merge_defaults(data.setdefault(k, {}), v)
        else:
            data.setdefault(k, v)


def create(data: dict) -> DictConfig:
    data = dict(data)
    method = (data.get("experiment") or {}).get("method")
    if method in DEFAULTS:
        _merge_defaults(data, DEFAULTS[method])
    return DictConfig(data)


def load(path: str) -> DictConfig:
    with open(path) as f:
        return create(yaml.safe_load(f))


class OmegaConf:
    """Name-compatible facade: ``OmegaConf.load`` / ``OmegaConf.to_container`` / ``OmegaConf.create``."""

    load = staticmethod(load)
    create = staticmethod(create)
    to_container = staticmethod(to_container)
