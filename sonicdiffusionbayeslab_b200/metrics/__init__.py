from .metrics import *  # noqa: F401,F403
