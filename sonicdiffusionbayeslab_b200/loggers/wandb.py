"""Logger facade of the drivers (/root/reference/src/loggers/wandb.py:42-91).

wandb is a network service and out of scope; unlike the reference,
``wandb_enable=False`` really is a null logger here (the reference dereferences ``wandb_logger``
regardless, loggers/wandb.py:72-91).  With ``wandb_enable=True`` and no ``$WANDB_KEY`` the logger
degrades to null with a warning instead of crashing.
"""
from __future__ import annotations

import logging
import os
from collections import defaultdict


class Logger:
    def __init__(self, config: dict, wandb_enable=True, project_name=None, run_name=None, run_id=None):
        self.logger = logging.getLogger("sonic")
        self.run = None
        if wandb_enable and os.environ.get("WANDB_KEY"):
            if project_name is None or run_name is None:
                raise ValueError("project_name and run_name are required when wandb is enabled")
            try:
                import wandb

                wandb.login(key=os.environ["WANDB_KEY"].strip())
                self.run = wandb.init(id=run_id or wandb.util.generate_id(), project=project_name, name=run_name,
                                      config=config or {}, resume="allow")
            except Exception as e:  # offline box: keep running
                self.logger.warning("wandb unavailable (%s): logging disabled", e)
        elif wandb_enable:
            self.logger.warning("wandb_enable is set but $WANDB_KEY is not: logging disabled")
        self.wandb_enable = self.run is not None
        self.losses_history = defaultdict(list)
        self.metrics_history = defaultdict(list)
        self.tables = {}

    def log_metrics(self, metrics: dict, step: int):
        for k, v in metrics.items():
            self.metrics_history[k].append((step, v))
        if self.run is not None:
            self.run.log({f"Metrics/{k}": v for k, v in metrics.items()}, step=step)

    def log_metrics_into_table(self, metrics: dict, name_table: str):
        self.tables[name_table] = {k: list(v) for k, v in metrics.items()}
        if self.run is not None:
            import pandas as pd
            import wandb

            self.run.log({name_table: wandb.Table(dataframe=pd.DataFrame.from_dict(metrics, orient="columns"))})

    def log_batch_of_images(self, images, name_images: str, captions=None):
        if self.run is not None:
            import wandb

            caps = captions or [None] * len(images)
            self.run.log({name_images: [wandb.Image(i, caption=c) for i, c in zip(images, caps)]})
