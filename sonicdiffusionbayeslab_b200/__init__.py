"""B200-native Stable Diffusion sampling engine with the plugin surface of SonicDiffusionBayesLab.

Importing the package registers every plugin (the reference does this in src/__init__.py:1-5).
Nothing here imports ``oracle/`` and nothing computes on the CPU: the hot path needs libsonic.so
(hand-written sm_100a kernels) and a CUDA device, and fails loudly otherwise.
"""
from . import schedulers  # noqa: F401
from . import models  # noqa: F401
from . import metrics  # noqa: F401
from . import experiments  # noqa: F401
from .registry import methods_registry, metrics_registry, models_registry, schedulers_registry  # noqa: F401
