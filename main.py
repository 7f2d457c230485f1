"""``python main.py --config dpm_solver_config.yaml`` -- same entry point as the reference
(/root/reference/main.py:10-24): load ./configs/<file>, seed, run the registered method."""
import argparse
import os

try:                                        # the real OmegaConf if it is installed ...
    from omegaconf import OmegaConf
except ImportError:                         # ... else the engine's OmegaConf-subset loader
    from sonicdiffusionbayeslab_b200.config import OmegaConf

from src.registry import methods_registry
from src.utils.model_utils import setup_seed


def main(config_file):
    from sonicdiffusionbayeslab_b200 import config as cfglib

    path = config_file if os.path.isfile(config_file) else os.path.join("./configs", config_file)
    config = OmegaConf.load(path)
    if not isinstance(config, cfglib.DictConfig):           # normalise + supply the keys the YAMLs omit
        config = cfglib.create(OmegaConf.to_container(config, resolve=True))
    setup_seed(config.experiment.get("seed", 29))
    methods_registry[config.experiment.method](config).run_experiment()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Sonic Diffusion (B200 engine)")
    parser.add_argument("--config", type=str, default="config.yaml", help="file under ./configs or a path")
    main(parser.parse_args().config)
