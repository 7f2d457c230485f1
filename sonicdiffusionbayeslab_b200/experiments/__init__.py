"""Experiment drivers ("methods" registry); importing the package registers them."""
from .base_experiment import BaseMethod  # noqa: F401
from .methods import *  # noqa: F401,F403
