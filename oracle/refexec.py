"""Execute the REFERENCE'S OWN SOURCE FILES over the oracle's restated third-party layer.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the pin of the oracle: the reference
cannot be imported as a package here (``import src`` dies on ``diffusers`` / ``omegaconf`` /
``torchmetrics``), but the files that hold the reference's OWN hot-path logic are plain Python on
top of a small third-party surface.  ``load()`` compiles those files *where they lie* under
``/root/reference`` (nothing is copied into the repo) with stub modules standing in for that
surface, so every line of

  * /root/reference/src/schedulers.py:12-197   (``DPMSolverScheduler.convert_model_output`` / ``.step``)
  * /root/reference/src/models.py:21-335       (``StableDiffusionModel.call``: the denoising loop)
  * /root/reference/src/models.py:338-730      (two-scheduler call + ``switch_timestamp``)
  * /root/reference/src/models.py:733-1135     (interleaved schedulers)
  * /root/reference/src/models.py:1138-1467    (skip-timesteps loop)
  * /root/reference/src/utils/class_registry.py, src/registry.py (plugin registry)

runs unmodified.  What the stubs provide -- and what therefore stays a *restatement* -- is the
diffusers layer underneath: ``DPMSolverMultistepScheduler`` / ``DDIMScheduler`` / ``LCMScheduler``
(= oracle/schedulers.py), ``StableDiffusionPipeline`` (the handful of helpers the loops call:
``check_inputs``, ``encode_prompt``, ``prepare_latents``, ``prepare_extra_step_kwargs``,
``progress_bar``, ``image_processor.postprocess``), ``retrieve_timesteps``, ``rescale_noise_cfg``,
``randn_tensor``, and the UNet (oracle/unet.py).

``/root/reference`` does not exist on the GPU box: ``available()`` is False there and the tests use
the fixtures ``tests/golden/make_reference_pins.py`` wrote from these executions.
"""
from __future__ import annotations

import contextlib
import inspect
import os
import sys
import types

import torch

from . import schedulers as O

REF_ROOT = os.environ.get("SONIC_REFERENCE_ROOT", "/root/reference")


def available(root: str = REF_ROOT) -> bool:
    return os.path.isfile(os.path.join(root, "src", "schedulers.py"))


# --------------------------------------------------------------------------- third-party stand-ins
def _retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """diffusers 0.32.1 ``retrieve_timesteps`` (pipeline_stable_diffusion.py)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed.")
    if timesteps is not None:
        if "timesteps" not in set(inspect.signature(scheduler.set_timesteps).parameters.keys()):
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             "custom timestep schedules.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    elif sigmas is not None:
        raise ValueError("custom sigmas are not on the reference path")
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        timesteps = scheduler.timesteps
    return timesteps, num_inference_steps


def _rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    std_text = noise_pred_text.std(dim=list(range(1, noise_pred_text.ndim)), keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


class _UNetShim:
    """What the loops touch of ``pipe.unet``: ``.config`` and the keyword call of models.py:227-235."""

    def __init__(self, net, deepcache=None):
        self.net = net
        self.deepcache = deepcache
        self.timesteps_list = None
        c = net.config
        self.config = types.SimpleNamespace(sample_size=c.sample_size, in_channels=c.in_channels,
                                            time_cond_proj_dim=c.time_cond_proj_dim)

    def __call__(self, sample, timestep, encoder_hidden_states=None, timestep_cond=None,
                 cross_attention_kwargs=None, added_cond_kwargs=None, return_dict=False):
        return self.net(sample, timestep, encoder_hidden_states=encoder_hidden_states)


class _IdentityVae:
    """``vae.decode(z / scaling_factor)`` -> ``(z,)``: the pins record latents, not pixels."""

    config = types.SimpleNamespace(scaling_factor=1.0)

    def decode(self, z, return_dict=False, generator=None):
        return (z,)


class _ImageProcessor:
    def postprocess(self, image, output_type="pt", do_denormalize=None):
        return image                       # identity: keeps the final latents readable from ``.images``


class _PipelineOutput:
    def __init__(self, images, nsfw_content_detected=None):
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected


class _StableDiffusionPipeline:
    """The slice of diffusers' ``StableDiffusionPipeline`` the reference loops call (restated)."""

    def __init__(self, unet, scheduler=None, device="cpu"):
        self.unet = _UNetShim(unet)
        self.scheduler = scheduler
        self.vae = _IdentityVae()
        self.vae_scale_factor = 8
        self.image_processor = _ImageProcessor()
        self._execution_device = torch.device(device)
        self._guidance_scale = 7.5
        self._guidance_rescale = 0.0
        self._clip_skip = None
        self._cross_attention_kwargs = None
        self._interrupt = False
        self._num_timesteps = 0

    guidance_scale = property(lambda s: s._guidance_scale)
    guidance_rescale = property(lambda s: s._guidance_rescale)
    clip_skip = property(lambda s: s._clip_skip)
    cross_attention_kwargs = property(lambda s: s._cross_attention_kwargs)
    interrupt = property(lambda s: s._interrupt)
    num_timesteps = property(lambda s: s._num_timesteps)

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1 and self.unet.config.time_cond_proj_dim is None

    def check_inputs(self, *a, **k):
        pass

    def encode_prompt(self, prompt, device, num_images_per_prompt, do_cfg, negative_prompt=None, prompt_embeds=None,
                      negative_prompt_embeds=None, lora_scale=None, clip_skip=None):
        if prompt_embeds is None:
            raise ValueError("the pin harness passes prompt_embeds (models.py:45-47)")

        def per_image(e):                                  # diffusers encode_prompt: the n images of a prompt adjacent
            if e is None:
                return None
            b, seq, _ = e.shape
            return e.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, seq, -1)

        return per_image(prompt_embeds), per_image(negative_prompt_embeds)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, int(height) // self.vae_scale_factor,
                 int(width) // self.vae_scale_factor)
        if latents is None:
            latents = O.randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield types.SimpleNamespace(update=lambda *a, **k: None)

    def maybe_free_model_hooks(self):
        pass


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    mods = {
        "omegaconf": mod("omegaconf", MISSING="???"),
        "diffusers": mod("diffusers", DDIMScheduler=O.DDIMScheduler, DPMSolverMultistepScheduler=O.DPMSolverScheduler,
                         LCMScheduler=O.LCMScheduler),
        "diffusers.schedulers": mod("diffusers.schedulers"),
        "diffusers.schedulers.scheduling_utils": mod("diffusers.schedulers.scheduling_utils",
                                                     SchedulerOutput=SchedulerOutput),
        "diffusers.utils": mod("diffusers.utils", deprecate=lambda *a, **k: None),
        "diffusers.utils.torch_utils": mod("diffusers.utils.torch_utils", randn_tensor=O.randn_tensor),
        "diffusers.callbacks": mod("diffusers.callbacks", MultiPipelineCallbacks=type("MultiPipelineCallbacks", (), {}),
                                   PipelineCallback=type("PipelineCallback", (), {})),
        "diffusers.image_processor": mod("diffusers.image_processor", PipelineImageInput=object),
        "diffusers.pipelines": mod("diffusers.pipelines"),
        "diffusers.pipelines.stable_diffusion": mod("diffusers.pipelines.stable_diffusion",
                                                    StableDiffusionPipelineOutput=_PipelineOutput),
        "diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion": mod(
            "diffusers.pipelines.stable_diffusion.pipeline_stable_diffusion",
            StableDiffusionPipeline=_StableDiffusionPipeline, rescale_noise_cfg=_rescale_noise_cfg,
            retrieve_timesteps=_retrieve_timesteps),
    }
    return mods


_CACHE = {}

# for the ``++`` algorithm types ``convert_model_output`` returns ONE tensor
# (schedulers.py:61) that ``step`` unpacks as two (schedulers.py:127), so the reference as written raises
# for every batch size but 2 (where it silently splits the batch).  ``load(patch_c1=True)`` applies this
# ONE-line source patch in memory -- the evident intent: solver input = x0_pred, returned x0_pred = x0_pred --
# so the ``++`` configurations (BASELINE configs[1]) can run through the reference's own loop code too.
C1_LINE = "        model_output, x0_pred = self.convert_model_output(model_output, sample=sample)\n"
C1_PATCH = ("        model_output = self.convert_model_output(model_output, sample=sample)\n"
            "        model_output, x0_pred = model_output if isinstance(model_output, tuple) else "
            "(model_output, model_output)\n")


def load(root: str = REF_ROOT, patch_c1: bool = False):
    """Compile and run the reference files listed in the module docstring; returns a namespace with the
    reference's classes.  ``sys.modules`` is restored afterwards (the repo's own ``src`` alias package and
    any real third-party module are untouched)."""
    key = (root, patch_c1)
    if key in _CACHE:
        return _CACHE[key]
    if not available(root):
        raise FileNotFoundError(f"{root}/src/schedulers.py not found (the reference tree is absent on the GPU box)")
    names = ["src", "src.utils", "src.utils.class_registry", "src.registry", "src.schedulers", "src.models"]
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in list(stubs) + names}
    try:
        sys.modules.update(stubs)
        for n in ("src", "src.utils"):
            pkg = types.ModuleType(n)
            pkg.__path__ = []
            sys.modules[n] = pkg
        loaded = {}
        for n, rel in (("src.utils.class_registry", "src/utils/class_registry.py"), ("src.registry", "src/registry.py"),
                       ("src.schedulers", "src/schedulers.py"), ("src.models", "src/models.py")):
            path = os.path.join(root, rel)
            m = types.ModuleType(n)
            m.__file__ = path
            sys.modules[n] = m
            with open(path) as f:
                text = f.read()
            if patch_c1 and n == "src.schedulers":
                if text.count(C1_LINE) != 1:
                    raise RuntimeError("reference schedulers.py changed: the C-1 line was not found exactly once")
                text = text.replace(C1_LINE, C1_PATCH)
            exec(compile(text, path, "exec"), m.__dict__)
            loaded[n] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    S, M, R = loaded["src.schedulers"], loaded["src.models"], loaded["src.registry"]
    ns = types.SimpleNamespace(
        root=root, schedulers=S, models=M, registry=R, ClassRegistry=loaded["src.utils.class_registry"].ClassRegistry,
        DPMSolverScheduler=S.DPMSolverScheduler, DDIMSchedulerMy=S.DDIMSchedulerMy, LCMScheduler=S.LCMScheduler,
        StableDiffusionModel=M.StableDiffusionModel,
        StableDiffusionModelTwoSchedulers=M.StableDiffusionModelTwoSchedulers,
        StableDiffusionModelInterlivingSchedulers=getattr(M, "StableDiffusionModelInterlivingSchedulers", None),
        StableDiffusionModelSkipTimesteps=getattr(M, "StableDiffusionModelSkipTimesteps", None))
    ns.patch_c1 = patch_c1
    _CACHE[key] = ns
    return ns


def load_plugins(root: str = REF_ROOT, features=None):
    """The reference's metric plugins, dataset and driver helpers compiled where they lie:
    ``src/metrics/metrics.py`` (``TimeMetric`` :115-131, ``ClipScoreMetric.calc_metric`` :25-41),
    ``src/dataset/dataset.py``, ``src/utils/model_utils.py`` and the script function ``calc_clip_score.py:13-37``.
    Third-party surface, restated: a minimal
    ``torchmetrics.Metric`` (``add_state`` / ``reset`` to the defaults) and torchmetrics 1.6.1 ``CLIPScore``'s state
    arithmetic (``score += sum_i 100 cos_i``, ``n_samples += n``, ``compute = max(score / n, 0)``; SURVEY appendix
    A.5) over an injectable feature function ``features(images, text) -> (f_img, f_txt)`` standing in for the CLIP
    towers.  ``ImageReward`` and the FID base are inert stubs (out of scope)."""
    ns = load(root)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    class Metric(torch.nn.Module):
        def __init__(self, **kwargs):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default.clone()
            setattr(self, name, default.clone())

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

    class CLIPScore(Metric):
        def __init__(self, model_name_or_path="openai/clip-vit-large-patch14", **kwargs):
            super().__init__(**kwargs)
            self.model_name_or_path = model_name_or_path
            self.add_state("score", torch.tensor(0.0), dist_reduce_fx="sum")
            self.add_state("n_samples", torch.tensor(0, dtype=torch.long), dist_reduce_fx="sum")
            self.seen = []

        def update(self, images, text):
            fi, ft = features(images, text)
            self.seen.append((images, list(text) if not isinstance(text, str) else [text]))
            score = 100 * (fi * ft).sum(axis=-1)
            self.score += score.sum(0)
            self.n_samples += len(score)

        def compute(self):
            return torch.max(self.score / self.n_samples, torch.zeros_like(self.score))

    stubs = {
        "ImageReward": mod("ImageReward", load=lambda **k: None),
        "torchmetrics": mod("torchmetrics", Metric=Metric),
        "torchmetrics.image": mod("torchmetrics.image"),
        "torchmetrics.image.fid": mod("torchmetrics.image.fid", FrechetInceptionDistance=type(
            "FrechetInceptionDistance", (Metric,), {})),
        "torchmetrics.multimodal": mod("torchmetrics.multimodal"),
        "torchmetrics.multimodal.clip_score": mod("torchmetrics.multimodal.clip_score", CLIPScore=CLIPScore),
    }
    names = ["src", "src.utils", "src.registry", "src.utils.model_utils", "src.metrics", "src.metrics.metrics",
             "src.dataset", "src.dataset.dataset", "reference_calc_clip_score"]
    saved = {k: sys.modules.get(k) for k in list(stubs) + names}
    try:
        sys.modules.update(stubs)
        for n in ("src", "src.utils", "src.metrics", "src.dataset"):
            pkg = types.ModuleType(n)
            pkg.__path__ = []
            sys.modules[n] = pkg
        sys.modules["src.registry"] = ns.registry
        loaded = {}
        for n, rel in (("src.utils.model_utils", "src/utils/model_utils.py"),
                       ("src.metrics.metrics", "src/metrics/metrics.py"), ("src.dataset.dataset", "src/dataset/dataset.py"),
                       ("reference_calc_clip_score", "calc_clip_score.py")):
            path = os.path.join(root, rel)
            m = types.ModuleType(n)
            m.__file__ = path
            sys.modules[n] = m
            with open(path) as f:
                exec(compile(f.read(), path, "exec"), m.__dict__)
            loaded[n] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return types.SimpleNamespace(metrics=loaded["src.metrics.metrics"], dataset=loaded["src.dataset.dataset"],
                                 model_utils=loaded["src.utils.model_utils"], registry=ns.registry,
                                 calc_clip_score=loaded["reference_calc_clip_score"].calc_clip_score)


class Recorder:
    """``callback_on_step_end``: records the latents after every step (models.py:263-273) and, optionally,
    teacher-forces the next step from a given list."""

    def __init__(self):
        self.per_step, self.timesteps = [], []

    def __call__(self, pipe, i, t, kwargs):
        self.per_step.append(kwargs["latents"].clone())
        self.timesteps.append(int(t))
        return {}
