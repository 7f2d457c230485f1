"""Compatibility alias so code written against the reference layout (``from src.registry import
methods_registry``, ``from src.utils.model_utils import setup_seed`` -- /root/reference/main.py:6-7)
resolves to the B200 engine's plugins."""
import importlib
import sys

import sonicdiffusionbayeslab_b200 as _pkg

for _name in ("registry", "schedulers", "models", "metrics", "experiments", "utils", "utils.model_utils",
              "utils.class_registry", "dataset", "loggers", "config"):
    sys.modules[f"src.{_name}"] = importlib.import_module(f"sonicdiffusionbayeslab_b200.{_name}")
registry = sys.modules["src.registry"]
