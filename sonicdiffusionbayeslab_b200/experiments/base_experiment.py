"""Experiment base class: the caller of the hot path (/root/reference/src/experiments/base_experiment.py).

Same hook structure (``setup_exp_params`` .. ``setup_loggers``, ``generate``, ``validate``,
``run_experiment``) and the same call into the model plugin
(``model(prompts, num_inference_steps=, guidance_scale=, generator=, output_type="pt")``), so the
method classes read like the reference's.  Differences, all forced by the deployment or by reference
defects: bf16 instead of fp16 (C-11), synthetic prompts when the COCO files
are absent, null logger offline (C-14), true image counts in the time metric (C-9), prompts sharded
across ranks when launched under torchrun (whole batches per rank, the shared generator replayed over the
batches of other ranks so every image equals the single-process run's) with the CLIP-score states
all-reduced over NCCL.
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from collections import defaultdict

import torch
from torch.utils.data import DataLoader, Subset

from .. import config as cfglib
from .. import dist as D
from ..dataset import ImageDatasetWithPrompts, SyntheticPromptDataset
from ..loggers import Logger
from ..registry import metrics_registry, models_registry, schedulers_registry
from ..utils.model_utils import save_image, save_table, to_pil_image


class BaseMethod(ABC):
    def __init__(self, config):
        self.config = config
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        if self.device == "cuda":
            local = int(os.environ.get("LOCAL_RANK", 0)) % torch.cuda.device_count()   # ranks may share one GPU in tests
            torch.cuda.set_device(local)
            self.device = f"cuda:{local}"
        self.rank, self.world = D.init_from_env(torch.device(self.device) if "cuda" in self.device else None)
        self.setup_exp_params()
        self.setup_generator()
        self.setup_model()
        self.setup_scheduler()
        self.setup_dataset()
        self.setup_metrics()
        self.setup_loggers()

    @abstractmethod
    def run_experiment(self):
        pass

    def setup_exp_params(self):
        pass

    def setup_generator(self):
        self.generator = torch.Generator(device=self.device)
        self.generator.manual_seed(self.config.experiment.seed)

    def setup_model(self):
        name = self.config.model.model_name
        dtype = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32,
                 "float32": torch.float32}[self.config.model.get("dtype", "bf16")]
        self.model = models_registry[name].from_pretrained(
            self.config.model.pretrained_model, timestamps=self.config.model.get("timestamps", None),
            safety_checker=None, requires_safety_checker=False, torch_dtype=dtype,
            seed=self.config.experiment.get("seed", 29))
        self.model.to(self.device)

    def setup_scheduler(self, **kwargs):
        name = self.config.scheduler.scheduler_name
        # "" means "use the class default" (two_schedulers.py:27-38 passes "" for unset keys, C-3)
        kwargs = {k: v for k, v in kwargs.items() if v != "" and v is not None}
        self.model.scheduler = schedulers_registry[name].from_config(self.model.scheduler.config, **kwargs)

    def setup_dataset(self):
        ds = self.config.dataset
        self.image_size = ds.image_size
        if os.path.isdir(ds.img_dataset) and os.path.isfile(ds.prompts):
            from torchvision import transforms

            tf = transforms.Compose([transforms.Resize(self.image_size), transforms.CenterCrop(self.image_size),
                                     transforms.ToTensor()])
            self.test_dataset = ImageDatasetWithPrompts(ds.img_dataset, ds.prompts, transform=tf)
        else:
            self.test_dataset = SyntheticPromptDataset(n=ds.get("num_synthetic", 1000), image_size=self.image_size,
                                                       seed=self.config.experiment.get("seed", 29))

    def setup_metrics(self):
        q = self.config.quality_metrics
        self.metric_dict = defaultdict(list)
        self.clip_score_gen_metric = metrics_registry["clip_score"](model_name_or_path=q.clip_score.model_name_or_path)
        self.image_reward_metric = metrics_registry["image_reward"](model_name=q.image_reward.model_name,
                                                                     device=self.device)
        self.fid_metric = metrics_registry["fid"](feature=q.fid.feature, input_img_size=q.fid.input_img_size,
                                                  normalize=q.fid.normalize)
        self.time_metric = metrics_registry["time_metric"]()

    def setup_loggers(self):
        lg = self.config.logger
        self.logger = Logger(config=cfglib.to_container(self.config, resolve=True),
                             wandb_enable=lg.get("wandb_enable", True) and self.rank == 0,
                             project_name=lg.get("project_name", None), run_name=self.config.experiment_name,
                             run_id=lg.get("run_id", None))

    # ---------------------------------------------------------------- data sharding
    def _global_batches(self, batch_size):
        """Every batch of the sweep point in dataloader order (shuffle=False; ``inference.batch_count`` caps it,
        base_experiment.py:130-131) as (start, stop) item ranges."""
        n = len(self.test_dataset)
        count = self.config.inference.get("batch_count", None)
        if count is not None:
            n = min(n, count * batch_size)
        return [(s, min(n, s + batch_size)) for s in range(0, n, batch_size)]

    def _local_dataloader(self, batch_size):
        """This rank's contiguous block of whole batches, in dataloader order (shuffle=False)."""
        blocks = self._my_batches(batch_size)
        idx = [i for a, b in blocks for i in range(a, b)]
        return DataLoader(Subset(self.test_dataset, idx), batch_size=batch_size, shuffle=False)

    def _my_batches(self, batch_size):
        batches = self._global_batches(batch_size)
        return D.shard_batches(batches[-1][1] if batches else 0, batch_size, self.rank, self.world)

    # ---------------------------------------------------------------- generation (the hot-path caller)
    def generate(self, test_dataloader, steps, batch_size=1, guidance_scale=7.5, accumulate_x0=False, **call_kwargs):
        """base_experiment.py:122-163.  ONE generator serves every batch (and every noisy scheduler step) of the
        whole experiment (:51-53,149), so under sharding each rank walks the GLOBAL batch list in order: the
        batches it owns run on the engine, the others are *replayed* (``rng_only=True``: the pipeline draws the
        initial latents and the per-step noise it would have drawn and advances the generator, nothing else).
        Rank r's images are therefore bit-identical to images [its block] of the single-process run.
        ``x0_preds``: the LAST batch's list (base_experiment.py:145,163), or with ``accumulate_x0`` every batch's
        (the ``generate`` overrides of default_sd.py:26,49 and skip_steps_exp.py:37,62)."""
        gen_images_list, x0_preds, x0_all = [], [], []
        mine = set(self._my_batches(batch_size))
        local = iter(test_dataloader)
        for start, stop in self._global_batches(batch_size):
            kw = dict(num_inference_steps=steps) if steps is not None else {}
            if (start, stop) not in mine:
                self.model([""] * (stop - start), guidance_scale=guidance_scale, generator=self.generator,
                           output_type="pt", rng_only=True, **kw, **call_kwargs)
                continue
            batch = next(local)
            prompts = list(batch["prompt"])
            assert len(prompts) == stop - start
            out, inference_time, x0_preds = self.model(prompts, guidance_scale=guidance_scale,
                                                       generator=self.generator, output_type="pt", **kw, **call_kwargs)
            x0_all.extend(x0_preds)
            imgs = out.images.float().cpu()
            gen_images_list.extend(imgs[i] for i in range(imgs.shape[0]))
            self.time_metric.update(inference_time, len(prompts))
        return gen_images_list, (x0_all if accumulate_x0 else x0_preds)

    def validate(self, test_dataloader, gen_dataloader, name_images, name_table, additional_values=None,
                 x0_preds_dataloader=None):
        self.clip_score_gen_metric.to(self.device)
        # base_experiment.py:176-190: with an x0 loader the three loaders are zipped -- the SHORTEST one (usually the
        # x0 grids: one batch per ``batch_size`` denoising steps) bounds the batches that are validated, as there
        loaders = (test_dataloader, gen_dataloader) + ((x0_preds_dataloader,) if x0_preds_dataloader is not None else ())
        for idx, batch in enumerate(zip(*loaders)):
            input_batch, gen_images = batch[0], batch[1]
            x0_preds = batch[2] if x0_preds_dataloader is not None else None
            image_files, real_images, prompts = input_batch["image_file"], input_batch["image"], input_batch["prompt"]
            real_u8 = (real_images * 255).to(torch.uint8).cpu()
            gen_u8 = (gen_images * 255).to(torch.uint8).cpu()
            self.clip_score_gen_metric.update(gen_u8.to(self.device), list(prompts))
            self.image_reward_metric.update(real_u8, gen_u8, prompts)
            self.fid_metric.update(gen_u8, real=False)
            self.fid_metric.update(real_u8, real=True)
            if idx % self.config.logger.get("log_images_step", 1) == 0:
                k = self.config.experiment.get("number_save_images", 8)
                self.logger.log_batch_of_images(images=gen_u8[:k], name_images=name_images, captions=list(prompts)[:k])
            if x0_preds and idx % self.config.logger.get("log_x0_step", 1) == 0:          # base_experiment.py:216-223
                k = self.config.experiment.get("number_x0", 1)
                self.logger.log_batch_of_images(images=x0_preds[:k], name_images=name_images,
                                                captions=list(prompts)[:k])
            if self.config.logger.save:
                out_dir = self.config.logger.save_dir.format(experiment=self.config.experiment_name, args=name_images)
                for f, img in zip(image_files, gen_u8.unbind(0)):
                    save_image(out_dir, f, to_pil_image(img))
        self.clip_score_gen_metric.to(self.device)
        D.all_reduce_metric(self.clip_score_gen_metric)        # NCCL: (score_sum, n_samples)
        D.all_reduce_metric(self.time_metric.to(self.device))
        for k, v in (additional_values or {}).items():
            self.metric_dict[k].append(v)
        self.metric_dict["nfe"].append(self.model.num_timesteps)
        self.metric_dict["clip_score_gen_image"].append(self.clip_score_gen_metric.compute().item())
        self.metric_dict["image_reward"].append(self.image_reward_metric.compute().item())
        self.metric_dict["fid"].append(self.fid_metric.compute().item())
        self.metric_dict["time_metric"].append(self.time_metric.compute().item())
        # a run on random-init stand-ins must be recognisable in its own metric table
        self.metric_dict["weights"].append(f"model:{getattr(self.model, 'weights_source', 'provided')},"
                                           f"clip:{getattr(self.clip_score_gen_metric.model, 'source', 'provided')}")
        if self.rank == 0:
            if self.config.logger.save:
                import pandas as pd

                out_dir = self.config.logger.save_dir.format(experiment=self.config.experiment_name, args=name_images)
                save_table(out_dir, "metrics", pd.DataFrame.from_dict(self.metric_dict, orient="columns"))
            self.logger.log_metrics_into_table(metrics=self.metric_dict, name_table=name_table)
        for m in (self.fid_metric, self.clip_score_gen_metric, self.image_reward_metric, self.time_metric):
            m.reset()

    # ---------------------------------------------------------------- shared sweep skeleton
    def collate_grid(self, batch):
        """base_experiment.py:274-284: the x0 predictions of ``batch_size`` consecutive denoising steps as one grid."""
        from torchvision.utils import make_grid

        return [make_grid(torch.stack(list(images)), nrow=8, normalize=True, padding=2) for images in zip(*batch)]

    def _new_table(self):
        """Every driver starts its metric table afresh in ``run_experiment`` (``self.metric_dict = defaultdict(list)``,
        e.g. ddim.py:29; deep_cache.py:39 once per cache interval)."""
        self.metric_dict = defaultdict(list)

    def _sweep_point(self, batch_size, steps, name_images, guidance_scale=7.5, additional_values=None,
                     x0_log_name=None, x0_grids=False, **call_kwargs):
        loader = self._local_dataloader(batch_size)
        self.model.to(self.device)
        gen_images, x0_preds = self.generate(loader, steps, batch_size, guidance_scale=guidance_scale,
                                             accumulate_x0=x0_log_name is not None, **call_kwargs)
        self.model.to("cpu")
        if x0_log_name is not None:                       # default_sd.py:89-92 / skip_steps_exp.py:117-120
            self.logger.log_batch_of_images(images=x0_preds, name_images=x0_log_name)
        gen_loader = DataLoader(gen_images, batch_size=batch_size, shuffle=False)
        x0_loader = None
        if x0_grids and self.config.experiment_params.get("use_x0", False):              # e.g. ddim.py:41-56
            x0_loader = DataLoader(x0_preds, batch_size=batch_size, shuffle=False, collate_fn=self.collate_grid)
        self.validate(loader, gen_loader, name_images=name_images, name_table=f"{self.config.experiment_name}",
                      additional_values=additional_values, x0_preds_dataloader=x0_loader)
        return gen_images, x0_preds
