"""Plugin registry with the contract of /root/reference/src/utils/class_registry.py:8-68.

``reg.add_to_registry(name, arg_keys=None, stop_args=(...))`` is a class decorator that stores
the class under ``reg.classes[name]`` and a dataclass synthesised from the signature of
``cls.__init__`` under ``reg.args[name]``; ``reg[name]`` returns the class.  Required
parameters get OmegaConf's ``MISSING`` marker ("???"), ``None`` defaults become ``Optional[Any]``
fields, other defaults become fields typed by the default's type.
"""
from __future__ import annotations

import dataclasses
import inspect
import typing

MISSING = "???"          # omegaconf.MISSING


def _field_for(name: str, param: inspect.Parameter):
    if param.default is inspect.Parameter.empty:
        return (name, typing.Any, MISSING)
    if param.default is None:
        return (name, typing.Optional[typing.Any], None)
    return (name, type(param.default), dataclasses.field(default=param.default))


class ClassRegistry:
    def __init__(self):
        self.classes: dict = {}
        self.args: dict = {}
        self.arg_keys = None

    def __getitem__(self, item):
        return self.classes[item]

    def __contains__(self, item):
        return item in self.classes

    def make_dataclass_from_init(self, func, name, arg_keys, stop_args):
        fields = [_field_for(k, v) for k, v in inspect.signature(func).parameters.items() if k not in stop_args]
        if not arg_keys:
            return dataclasses.make_dataclass(name, fields)
        self.arg_keys = arg_keys
        per_key = {key: dataclasses.make_dataclass(key, fields) for key in arg_keys}
        return dataclasses.make_dataclass(
            name, [(k, v, dataclasses.field(default_factory=v)) for k, v in per_key.items()])

    def make_dataclass_from_classes(self, name):
        return dataclasses.make_dataclass(
            name, [(k, v, dataclasses.field(default_factory=v)) for k, v in self.classes.items()])

    def make_dataclass_from_args(self, name):
        return dataclasses.make_dataclass(
            name, [(k, v, dataclasses.field(default_factory=v)) for k, v in self.args.items()])

    def add_to_registry(self, name, arg_keys=None, stop_args=("self", "args", "kwargs")):
        def register(cls):
            self.classes[name] = cls
            self.args[name] = self.make_dataclass_from_init(cls.__init__, name, arg_keys, stop_args)
            return cls

        return register
