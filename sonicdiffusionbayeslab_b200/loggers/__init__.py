from .wandb import Logger  # noqa: F401
