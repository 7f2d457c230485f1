"""Experiment-driver pins: the reference's OWN drivers (/root/reference/src/experiments/*.py, executed where they lie
by ``load_reference_drivers``) and the product's drivers (sonicdiffusionbayeslab_b200/experiments) are run over the same
RECORDING fake backend -- model plugin, metric plugins, DeepCache helper, logger, dataset -- and must produce the same
event log: which scheduler is built with which keys, which LoRA / DeepCache calls are made, every call into the model
plugin with its keyword arguments, every sweep point's image-log name, and every metric table with its extra columns.

TEST INFRASTRUCTURE ONLY.  Shared by tests/golden/make_reference_pins.py (writes
``tests/golden/reference_driver_events.json`` from the reference side) and tests/test_reference_drivers_cpu.py.
"""
from __future__ import annotations

import os
import sys
import types

import torch

from oracle import refexec
from oracle import schedulers as O

SD15 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
            trained_betas=None, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1, clip_sample=False,
            prediction_type="epsilon", timestep_spacing="leading")
N_ITEMS, BATCH = 5, 2                                   # 3 batches, the last one ragged


def _base(method, model_name, scheduler, params, **extra):
    cfg = {
        "experiment_name": f"pin {method}",
        "experiment": {"method": method, "seed": 29, "number_save_images": 2},
        "model": {"model_name": model_name, "pretrained_model": "runwayml/stable-diffusion-v1-5"},
        "scheduler": scheduler,
        "dataset": {"img_dataset": "/nonexistent/images", "prompts": "/nonexistent/prompts.json", "image_size": 8,
                    "num_synthetic": N_ITEMS},
        "quality_metrics": {"clip_score": {"model_name_or_path": "openai/clip-vit-base-patch16"},
                            "image_reward": {"model_name": "ImageReward-v1.0"},
                            "fid": {"feature": 64, "input_img_size": 8, "normalize": False}},
        "logger": {"wandb_enable": False, "project_name": "p", "log_images_step": 2, "save": False,
                   "save_dir": "/nonexistent/{experiment}/{args}/"},
        "inference": {"batch_size": BATCH},
        "experiment_params": params,
    }
    cfg.update(extra)
    return cfg


DPM = {"solver_order": 2, "algorithm_type": "dpmsolver++", "final_sigmas_type": "zero"}
SD = "stable_diffusion_model"
DRIVER_CASES = {
    "ddim": _base("ddim", SD, {"scheduler_name": "ddim_scheduler"}, {"num_inference_steps": [4, 6]}),
    "ddim_use_x0": _base("ddim", SD, {"scheduler_name": "ddim_scheduler"}, {"num_inference_steps": [4], "use_x0": True}),
    "dpm_solver": _base("dpm_solver", SD, {"scheduler_name": "dpm_solver_scheduler"},
                        {"num_inference_steps": [5, 7], **DPM}),
    "dpm_solver_batch_count": _base("dpm_solver", SD, {"scheduler_name": "dpm_solver_scheduler"},
                                    {"num_inference_steps": [5], **DPM, "solver_order": 3},
                                    inference={"batch_size": BATCH, "batch_count": 2}),
    "consistency_model": _base("consistency_model", SD, {"scheduler_name": "lcm_scheduler"},
                               {"num_inference_steps": [2, 4], "guidance_scale": 0.0,
                                "adapter_id": "latent-consistency/lcm-lora-sdv1-5"}),
    "deep_cache": _base("deep_cache", SD, {}, {"num_inference_steps": [6, 9], "cache_interval": [2, 5],
                                               "cache_branch_id": 1}),
    "default": _base("default", SD, {}, {"num_inference_steps": [5, 8]}),
    "two_schedulers": _base("two_schedulers", "stable_diffusion_model_two_schedulers",
                            {"scheduler_first": "ddim_scheduler", "scheduler_second": "dpm_solver_scheduler"},
                            {"num_inference_steps_first": [10, 20], "num_inference_steps_second": [10, 20],
                             "num_step_switch": [3, 10], "type_switch": "closest", "solver_order": 2,
                             "first_algorithm_type": "dpmsolver++", "first_final_sigmas_type": "zero",
                             "first_order_solver": 3, "second_algorithm_type": "dpmsolver",
                             "second_final_sigmas_type": "sigma_min", "second_order_solver": 1}),
    "two_schedulers_use_x0": _base("two_schedulers", "stable_diffusion_model_two_schedulers",
                                   {"scheduler_first": "dpm_solver_scheduler", "scheduler_second": "dpm_solver_scheduler"},
                                   {"num_inference_steps_first": [10], "num_inference_steps_second": [10],
                                    "num_step_switch": [4], "type_switch": "left_closest", "use_x0": True,
                                    "first_algorithm_type": "dpmsolver++", "first_final_sigmas_type": "zero",
                                    "second_algorithm_type": "dpmsolver++", "second_final_sigmas_type": "zero"}),
    "skip_steps": _base("skip_steps", "stable_diffusion_model_skip_timesteps", {"scheduler_name": "dpm_solver_scheduler"},
                        {"num_inference_steps": [8, 10], "skip_steps": [[2, 5], [1, 2, 7]], **DPM}),
    "interliving_schedulers": _base("interliving_schedulers", "stable_diffusion_model_interliving_schedulers",
                                    {"scheduler_main": "dpm_solver_scheduler", "scheduler_inter": "ddim_scheduler"},
                                    {"num_inference_steps_first": [10, 12], "interliving_steps": [[1, 3], [0, 2, 4]],
                                     "main_order_solver": 2, "main_algorithm_type": "dpmsolver++",
                                     "main_final_sigmas_type": "zero", "inter_order_solver": "",
                                     "inter_algorithm_type": "", "inter_final_sigmas_type": ""}),
}
# the shipped two_schedulers_config.yaml leaves the algorithm keys unset: the reference passes "" to from_config and
# the DPM constructor refuses it (SURVEY C-3) -- pinned as "the reference raises", the product treats "" as the default
RAISING_CASES = {
    "two_schedulers_shipped_keys": _base("two_schedulers", "stable_diffusion_model_two_schedulers",
                                         {"scheduler_first": "ddim_scheduler", "scheduler_second": "dpm_solver_scheduler"},
                                         {"num_inference_steps_first": [10], "num_inference_steps_second": [10],
                                          "num_step_switch": [3], "type_switch": "closest", "solver_order": 2}),
}


# ----------------------------------------------------------------------------------------- the recording backend
def _plain(v):
    """JSON-able, order-stable rendering of an argument value."""
    if isinstance(v, (str, int, float, bool)) or v is None:
        return v
    if isinstance(v, torch.Tensor):
        return f"tensor{tuple(v.shape)}"
    if isinstance(v, dict) or hasattr(v, "items"):
        return {str(k): _plain(x) for k, x in v.items()}
    try:
        return [_plain(x) for x in v]
    except TypeError:
        return type(v).__name__


SCHED_KEYS = ("solver_order", "algorithm_type", "final_sigmas_type", "solver_type", "steps_offset", "timestep_spacing",
              "beta_schedule", "clip_sample")


class Backend:
    """One event log + the fake plugins writing into it."""

    def __init__(self, default_scheduler):
        self.events = []
        log = self.events

        def sched_event(slot, s):
            cfg = getattr(s, "config", {})
            log.append(["scheduler", slot, type(s).__name__, {k: _plain(cfg[k]) for k in SCHED_KEYS if k in cfg}])

        class FakeModel:
            num_timesteps = 0

            def __init__(self, name):
                self.__dict__["_slots"] = {"scheduler": default_scheduler()}
                self.name = name
                self.unet = types.SimpleNamespace(config=types.SimpleNamespace(sample_size=8, in_channels=4))

            @classmethod
            def make_class(cls, name):
                def from_pretrained(klass, pretrained, timestamps=None, safety_checker=None,
                                    requires_safety_checker=False, **kw):
                    log.append(["from_pretrained", name, pretrained, _plain(timestamps), _plain(safety_checker),
                                requires_safety_checker])
                    return klass(name)

                return type(f"Fake_{name}", (cls,), {"from_pretrained": classmethod(from_pretrained)})

            def __setattr__(self, k, v):
                if k in ("scheduler", "scheduler_first", "scheduler_second", "scheduler_main", "scheduler_inter"):
                    self._slots[k] = v
                    sched_event(k, v)
                else:
                    object.__setattr__(self, k, v)

            def __getattr__(self, k):
                slots = self.__dict__.get("_slots", {})
                if k in slots:
                    return slots[k]
                raise AttributeError(k)

            def to(self, device):
                return self

            def load_lora_weights(self, adapter, **kw):
                log.append(["load_lora_weights", adapter, _plain(kw)])

            def fuse_lora(self, **kw):
                log.append(["fuse_lora", _plain(kw)])

            def __call__(self, prompts, generator=None, **kw):
                n = len(prompts)
                kw.pop("rng_only", None)
                steps = kw.get("num_inference_steps") or kw.get("num_inference_steps_first") or 0
                object.__setattr__(self, "num_timesteps", int(steps))
                log.append(["call", n, {k: _plain(v) for k, v in sorted(kw.items())},
                            isinstance(generator, torch.Generator)])
                g = torch.Generator().manual_seed(n)
                imgs = torch.rand(n, 3, 8, 8, generator=g)
                x0 = [torch.rand(1, 3, 8, 8, generator=g) for _ in range(3)]     # x0_pred[0] of three denoising steps
                return types.SimpleNamespace(images=imgs), 0.25, x0

        class FakeMetric:
            def __init__(self, *a, **kw):
                self.kind = "?"
                self.n = 0

            def to(self, device):
                return self

            def update(self, *a, **kw):
                self.n += 1
                if self.kind == "time_metric":
                    log.append(["time_update", _plain(a[1]) if len(a) > 1 else None])

            def compute(self):
                return torch.tensor(float(self.n))

            def reset(self):
                self.n = 0

        def metric_class(kind):
            def init(self, *a, **kw):
                FakeMetric.__init__(self)
                self.kind = kind
                self.model = types.SimpleNamespace(source="provided")
                log.append(["metric", kind, {k: _plain(v) for k, v in sorted(kw.items()) if k != "device"}])

            return type(f"Fake_{kind}", (FakeMetric,), {"__init__": init})

        class FakeLogger:
            def __init__(self, config=None, wandb_enable=True, project_name=None, run_name=None, run_id=None):
                log.append(["logger", bool(wandb_enable), project_name, run_name])

            def log_batch_of_images(self, images, name_images, captions=None):
                log.append(["images", name_images, len(images)])

            def log_metrics_into_table(self, metrics, name_table):
                extra = {k: _plain(v) for k, v in metrics.items()
                         if k not in ("clip_score_gen_image", "image_reward", "fid", "time_metric", "weights")}
                log.append(["table", name_table, sorted(k for k in metrics if k != "weights"), extra])

        class FakeDeepCache:
            def __init__(self, pipe=None):
                log.append(["deepcache", "init", type(pipe).__name__.startswith("Fake_")])

            def set_params(self, **kw):
                log.append(["deepcache", "set_params", {k: _plain(v) for k, v in sorted(kw.items())}])

            def enable(self, *a):
                log.append(["deepcache", "enable"])

            def disable(self):
                log.append(["deepcache", "disable"])

        class FakeDataset(torch.utils.data.Dataset):
            def __init__(self, *a, **kw):
                pass

            def __len__(self):
                return N_ITEMS

            def __getitem__(self, i):
                return {"image_file": f"img_{i}.png", "image": torch.zeros(3, 8, 8), "prompt": f"prompt {i}"}

        self.FakeModel, self.metric_class, self.FakeLogger = FakeModel, metric_class, FakeLogger
        self.FakeDeepCache, self.FakeDataset = FakeDeepCache, FakeDataset


MODEL_NAMES = ("stable_diffusion_model", "stable_diffusion_model_two_schedulers",
               "stable_diffusion_model_interliving_schedulers", "stable_diffusion_model_skip_timesteps")
METRIC_NAMES = ("clip_score", "image_reward", "fid", "time_metric")


def _config(case):
    from sonicdiffusionbayeslab_b200 import config as cfglib

    return cfglib.create(case)


# ----------------------------------------------------------------------------------------- reference side
_DRIVER_FILES = ("base_experiment", "ddim", "dpm_solver", "consistency_model", "deep_cache", "default_sd",
                 "two_schedulers", "skip_steps_exp", "interliving_exp")


def run_reference_driver(case):
    """One case through the reference's own driver source.  Returns the event log."""
    from sonicdiffusionbayeslab_b200 import config as cfglib

    ns = refexec.load(patch_c1=True)
    be = Backend(lambda: O.PNDMScheduler.from_config(SD15))
    reg = ns.registry
    saved = {r: (dict(getattr(reg, r).classes), dict(getattr(reg, r).args))
             for r in ("models_registry", "metrics_registry", "methods_registry")}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    stubs = {
        "omegaconf": mod("omegaconf", OmegaConf=types.SimpleNamespace(
            to_container=lambda cfg, resolve=True: cfglib.to_container(cfg, resolve=resolve))),
        "DeepCache": mod("DeepCache", DeepCacheSDHelper=be.FakeDeepCache),
        "src": mod("src"), "src.registry": reg, "src.schedulers": ns.schedulers, "src.models": ns.models,
        "src.dataset": mod("src.dataset"), "src.dataset.dataset": mod("src.dataset.dataset",
                                                                        ImageDatasetWithPrompts=be.FakeDataset),
        "src.loggers": mod("src.loggers"), "src.loggers.wandb": mod("src.loggers.wandb", Logger=be.FakeLogger),
        "src.utils": mod("src.utils"),
        "src.utils.model_utils": mod("src.utils.model_utils", save_image=lambda *a, **k: None,
                                     save_table=lambda *a, **k: None, to_pil_image=lambda t: t),
        "src.experiments": mod("src.experiments"),
    }
    for m in ("src", "src.dataset", "src.loggers", "src.utils", "src.experiments"):
        stubs[m].__path__ = []
    names = list(stubs) + [f"src.experiments.{f}" for f in _DRIVER_FILES]
    keep = {k: sys.modules.get(k) for k in names}
    try:
        sys.modules.update(stubs)
        for name in MODEL_NAMES:
            reg.models_registry.classes[name] = be.FakeModel.make_class(name)
        for kind in METRIC_NAMES:
            reg.metrics_registry.classes[kind] = be.metric_class(kind)
        for f in _DRIVER_FILES:
            path = os.path.join(ns.root, "src", "experiments", f + ".py")
            m = types.ModuleType(f"src.experiments.{f}")
            m.__file__ = path
            sys.modules[m.__name__] = m
            with open(path) as fh:
                exec(compile(fh.read(), path, "exec"), m.__dict__)
        cfg = _config(case)
        method = reg.methods_registry[cfg.experiment.method](cfg)
        be.events.append(["generator", method.generator.initial_seed()])
        method.run_experiment()
    finally:
        for r, (classes, args) in saved.items():
            getattr(reg, r).classes.clear()
            getattr(reg, r).classes.update(classes)
            getattr(reg, r).args.clear()
            getattr(reg, r).args.update(args)
        for k, v in keep.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return be.events


# ----------------------------------------------------------------------------------------- product side
def run_product_driver(case, monkeypatch):
    """The same case through sonicdiffusionbayeslab_b200.experiments over the same fake backend."""
    import sonicdiffusionbayeslab_b200  # noqa: F401  (registration)
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.experiments import base_experiment as BE
    from sonicdiffusionbayeslab_b200.experiments import methods as ME
    from sonicdiffusionbayeslab_b200.registry import methods_registry, metrics_registry, models_registry

    be = Backend(lambda: S.PNDMScheduler.from_config(SD15))
    for name in MODEL_NAMES:
        monkeypatch.setitem(models_registry.classes, name, be.FakeModel.make_class(name))
    for kind in METRIC_NAMES:
        monkeypatch.setitem(metrics_registry.classes, kind, be.metric_class(kind))
    monkeypatch.setattr(BE, "Logger", be.FakeLogger)
    monkeypatch.setattr(BE, "SyntheticPromptDataset", be.FakeDataset)
    monkeypatch.setattr(ME, "DeepCacheSDHelper", be.FakeDeepCache)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    cfg = _config(case)
    method = methods_registry[cfg.experiment.method](cfg)
    be.events.append(["generator", method.generator.initial_seed()])
    method.run_experiment()
    return be.events


def normalise(events):
    """The deliberate differences, removed before the logs are compared:
      * C-9: the reference counts a short last batch as ``batch_size`` images in the time metric
        (base_experiment.py:161), the product counts the images it made -> ``time_update`` is checked separately;
      * C-3: the reference forwards unset solver keys as ``""`` (they end up as inert hidden config entries of a
        scheduler that has no such field, or make the DPM constructor raise -- RAISING_CASES); the product drops them;
      * the product's logger is silent off rank 0 / offline -> ``wandb_enable`` is not compared."""
    out = []
    for e in events:
        e = list(e)
        if e[0] == "time_update":
            continue
        if e[0] == "logger":
            e = ["logger", e[2], e[3]]
        if e[0] == "scheduler":
            e[3] = {k: v for k, v in e[3].items() if v != ""}
        out.append(e)
    return out


def time_updates(events):
    return [e[1] for e in events if e[0] == "time_update"]
