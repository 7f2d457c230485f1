"""Pipeline plugins: the B200 sampling engine behind the reference's ``models_registry`` names.

``stable_diffusion_model``                 <- /root/reference/src/models.py:21-335
``stable_diffusion_model_two_schedulers``  <- /root/reference/src/models.py:338-730
``stable_diffusion_model_skip_timesteps``  <- /root/reference/src/models.py:1138-1467 (step mask)

Same call contract as the reference pipelines: ``model(prompts, num_inference_steps=, guidance_scale=,
generator=, output_type="pt", latents=, prompt_embeds=, negative_prompt_embeds=, ...)`` returns
``(output_with_.images, execution_time_seconds, x0_preds)``; ``execution_time`` covers the denoising
loop only (models.py:208,284-285) but is measured with a device synchronise on both sides.

The loop body of models.py:211-261 becomes, per step, one replay of a native launch plan (the UNet,
``UNetEngine``) plus ONE fused kernel (CFG combine + scheduler update + x0 + history, through the
scheduler plugin's ``step_cfg``); latents stay resident in the engine's input buffer.  The
timestep / index schedule is computed on the host exactly as diffusers does (bit-exact).
"""
from __future__ import annotations

import inspect
import logging
import os
import time
import warnings
from types import SimpleNamespace

import numpy as np
import torch

from . import schedulers as S
from .registry import models_registry
from .text import load_tokenizer, make_text_encoder
from .unet_engine import PackedWeights, UNetArch, UNetEngine
from .unet_spec import random_unet_state_dict, validate_state_dict
from .vae_spec import VaeWeights, random_vae_state_dict, validate_vae_state_dict

log = logging.getLogger("sonicdiffusionbayeslab_b200")


def _loud(msg):
    """Random-init / synthetic stand-ins must never pass for a real run: warn on both channels."""
    log.warning(msg)
    warnings.warn(msg, RuntimeWarning, stacklevel=3)

# runwayml/stable-diffusion-v1-5 scheduler/scheduler_config.json + instantiated defaults
SD15_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    trained_betas=None, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1, clip_sample=False,
    prediction_type="epsilon", timestep_spacing="leading",
)


def load_stock_scheduler(model_dir=None):
    """The scheduler ``from_pretrained`` leaves in ``pipe.scheduler`` -- what the ``default`` / ``deep_cache`` methods
    run with (default_sd.py:15-16, deep_cache.py:17-18) and whose ``.config`` every other method feeds to
    ``from_config`` (base_experiment.py:69-72).  A local diffusers-layout directory is read like diffusers reads it
    (``scheduler/scheduler_config.json``: ``_class_name`` picks the class, the remaining keys are its config); without
    one it is SD-v1.5's PNDM.  A stock class this package has no fused step for keeps its CONFIG (so the
    ``from_config`` idiom sees the right betas / spacing / prediction type) on a PNDM stand-in, with a loud warning."""
    path = os.path.join(model_dir, "scheduler", "scheduler_config.json") if model_dir else None
    if not (path and os.path.isfile(path)):
        return S.PNDMScheduler.from_config(SD15_SCHEDULER_CONFIG)
    import json

    with open(path) as f:
        cfg = json.load(f)
    name = cfg.pop("_class_name", "PNDMScheduler")
    cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
    known = {"PNDMScheduler": S.PNDMScheduler, "DDIMScheduler": S.DDIMSchedulerMy,
             "DPMSolverMultistepScheduler": S.DPMSolverScheduler, "LCMScheduler": S.LCMScheduler}
    why = None
    if name in known:
        try:
            return known[name].from_config(cfg)
        except (NotImplementedError, ValueError) as e:      # an option of that class without a fused step (e.g. Karras sigmas)
            why = f"stock scheduler {name} with this configuration is not fused here ({e})"
    _loud(f"{path}: {why or f'stock scheduler class {name!r} has no fused step here'}; its config is kept for the "
          "`from_config` idiom, but the `default` / `deep_cache` methods (which run the stock scheduler itself) "
          "would step with PNDM instead")
    try:
        return S.PNDMScheduler.from_config({**SD15_SCHEDULER_CONFIG, **cfg})
    except (NotImplementedError, ValueError):             # e.g. a spacing PLMS is not fused for: config only
        sched = S.PNDMScheduler.from_config(SD15_SCHEDULER_CONFIG)
        sched.config.update(cfg)
        return sched


class PipelineOutput(SimpleNamespace):
    """``StableDiffusionPipelineOutput`` stand-in: ``.images`` and ``.nsfw_content_detected``."""


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, **kwargs):
    """diffusers' helper as used at models.py:167-169,487-494."""
    if timesteps is not None:
        if "timesteps" not in inspect.signature(scheduler.set_timesteps).parameters:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not "
                             "support custom timestep schedules.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        return scheduler.timesteps, len(scheduler.timesteps)
    scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, num_inference_steps


def postprocess_images(image01: torch.Tensor, output_type: str):
    """diffusers ``VaeImageProcessor.postprocess`` on already-denormalised [0, 1] tensors (models.py:313-315):
    ``"pt"`` the tensor, ``"np"`` float32 NHWC numpy, ``"pil"`` a list of PIL images (``(x * 255).round()`` uint8)."""
    if output_type == "pt":
        return image01
    arr = image01.cpu().permute(0, 2, 3, 1).float().numpy()
    if output_type == "np":
        return arr
    if output_type == "pil":
        from PIL import Image

        u8 = (arr * 255).round().astype("uint8")
        return [Image.fromarray(a.squeeze(-1), mode="L") if a.shape[-1] == 1 else Image.fromarray(a) for a in u8]
    raise ValueError(f"output_type {output_type!r}: expected one of 'latent', 'pt', 'np', 'pil'")


def _load_safetensors(path):
    from safetensors.torch import load_file

    return load_file(path)


class _PipelineBase:
    """State shared by the pipeline variants: weights, engines, text/VAE modules, DeepCache flags."""

    def __init__(self, unet_state_dict, vae, text_encoder, tokenizer, scheduler, *, arch=None,
                 torch_dtype=torch.bfloat16, latent_size=64, timestamps=None, seed=29):
        arch = arch or UNetArch()
        self.arch = arch
        self.dtype = torch_dtype
        self.latent_size = latent_size
        self.vae = vae
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.scheduler = scheduler
        self.timestamps = timestamps
        self.vae_scale_factor = 8
        self.device = torch.device("cpu")
        self._unet_sd = unet_state_dict
        self._weights = None
        self._engines = {}
        self._num_timesteps = 0
        self._deepcache = None                      # set by DeepCacheSDHelper.enable()
        self.use_cuda_graphs = True
        self.decode_x0_preds = True
        self.weights_source = "provided"            # or "random-init": reported in the experiments' metric tables
        self.unet = SimpleNamespace(config=SimpleNamespace(
            in_channels=arch.in_channels, sample_size=latent_size, time_cond_proj_dim=None))
        self.last_step_kinds = []                   # "full"/"cached" per UNet call of the last run

    # ---------------------------------------------------------------- construction
    @classmethod
    def from_pretrained(cls, pretrained_model, timestamps=None, safety_checker=None,
                        requires_safety_checker=False, torch_dtype=torch.bfloat16, seed=29, **kwargs):
        """``pretrained_model``: a local diffusers-layout directory (unet/, vae/, text_encoder/, tokenizer/) is
        loaded.  Any other id (e.g. "runwayml/stable-diffusion-v1-5": there is no network) yields seeded
        RANDOM-INIT weights of the SD-v1.5 architecture and a hash tokenizer -- the synthetic workload of
        BASELINE.json -- with a loud warning; ``SONIC_REQUIRE_WEIGHTS=1`` turns that into an error."""
        if torch_dtype == torch.float16:
            torch_dtype = torch.bfloat16            # the engine computes in bf16 (SURVEY C-11)
        arch = UNetArch()
        root = pretrained_model if isinstance(pretrained_model, str) and os.path.isdir(pretrained_model) else None
        sub = (lambda *p: os.path.join(root, *p)) if root else None
        random_parts = []
        unet_file = sub("unet", "diffusion_pytorch_model.safetensors") if root else None
        if unet_file and os.path.exists(unet_file):
            sd = _load_safetensors(unet_file)
            validate_state_dict(sd, arch)
        else:
            sd = random_unet_state_dict(seed, arch)
            random_parts.append("UNet")
        vae_file = sub("vae", "diffusion_pytorch_model.safetensors") if root else None
        if vae_file and os.path.exists(vae_file):
            vae_sd = validate_vae_state_dict(_load_safetensors(vae_file))
        else:
            vae_sd = random_vae_state_dict(seed)
            random_parts.append("VAE decoder")
        text_dir = sub("text_encoder") if root else None
        if not (text_dir and os.path.isdir(text_dir)):
            random_parts.append("text encoder")
        tok_dir = sub("tokenizer") if root else None
        if not (tok_dir and os.path.isdir(tok_dir)):
            random_parts.append("tokenizer (CRC32 hash tokenizer)")
        if random_parts:
            msg = (f"from_pretrained({pretrained_model!r}): no local weights for {', '.join(random_parts)} -- using "
                   f"seeded RANDOM-INIT stand-ins (seed {seed}); images and quality metrics of this run are "
                   "synthetic-workload numbers, not Stable Diffusion outputs")
            if os.environ.get("SONIC_REQUIRE_WEIGHTS") == "1":
                raise FileNotFoundError(msg)
            _loud(msg)
        text = make_text_encoder(seed, text_dir, dtype=torch_dtype)
        tok = load_tokenizer(tok_dir)
        sched = load_stock_scheduler(root)
        pipe = cls(sd, VaeWeights(vae_sd), text, tok, sched, arch=arch, torch_dtype=torch_dtype, timestamps=timestamps,
                   seed=seed)
        pipe.weights_source = "random-init" if random_parts else "provided"
        return pipe

    def to(self, device):
        """Weights are packed into HBM on the first move to a CUDA device and stay resident; moving
        "to cpu" (the reference bounces the model around every sweep point, ddim.py:31-33) keeps
        the engines alive and is a no-op."""
        dev = torch.device(device)
        if dev.type == "cuda":
            self.device = dev
        return self

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def guidance_rescale(self):
        return getattr(self, "_guidance_rescale", 0.0)

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1 and self.unet.config.time_cond_proj_dim is None

    def load_lora_weights(self, adapter, adapter_scale=1.0, **_):
        """LoRA merge source (consistency_model.py:20-21): a local ``.safetensors`` in either the diffusers/PEFT
        layout (``<module>.lora_A.weight`` / ``.lora_B.weight``) or the kohya layout the LCM-LoRA ships in
        (``lora_unet_<module with _>.lora_down.weight`` / ``.lora_up.weight`` / ``.alpha``).  A hub id cannot be
        fetched (no network): that is reported loudly and the base weights stay as they are."""
        self._lora = None
        if isinstance(adapter, str) and os.path.isfile(adapter):
            self._lora = (_load_safetensors(adapter), adapter_scale, adapter)
        else:
            _loud(f"load_lora_weights({adapter!r}): not a local file and there is no network -- NO adapter is "
                  "applied; the consistency-model method then runs the LCM scheduler on non-LCM weights")

    def fuse_lora(self, lora_scale=1.0):
        lora = getattr(self, "_lora", None)
        if not lora:
            return
        tensors, scale, path = lora
        flat = {k[: -len(".weight")].replace(".", "_"): k for k in self._unet_sd if k.endswith(".weight")}
        pairs = []                                   # (target key, down/A, up/B, alpha or None)
        for k in tensors:
            if k.endswith(".lora_A.weight"):
                base = k[: -len(".lora_A.weight")]
                target = base.replace("unet.", "", 1) + ".weight"
                pairs.append((target, tensors[k], tensors[base + ".lora_B.weight"], tensors.get(base + ".alpha")))
            elif k.endswith(".lora_down.weight"):
                base = k[: -len(".lora_down.weight")]
                if not base.startswith("lora_unet_"):
                    continue                          # text-encoder adapters (lora_te_*) are not applied
                target = flat.get(base[len("lora_unet_"):])
                pairs.append((target, tensors[k], tensors[base + ".lora_up.weight"], tensors.get(base + ".alpha")))
        merged = 0
        for target, A, B, alpha in pairs:
            if target is None or target not in self._unet_sd:
                continue
            A, B = A.float(), B.float()
            rank = A.shape[0]
            a = float(alpha) if alpha is not None else float(rank)
            delta = (B.flatten(1) @ A.flatten(1)) * (a / rank) * scale * lora_scale
            w = self._unet_sd[target].float()
            self._unet_sd[target] = w + delta.reshape(w.shape)
            merged += 1
        if merged == 0:
            raise ValueError(f"fuse_lora: none of the {len(tensors)} tensors in {path} matched a UNet weight "
                             "(expected diffusers `lora_A/lora_B` or kohya `lora_unet_*.lora_down/lora_up` keys)")
        self._weights = None
        self._engines = {}

    # ---------------------------------------------------------------- argument checks
    _callback_tensor_inputs = ["latents", "prompt_embeds", "negative_prompt_embeds"]

    def check_inputs(self, prompt, height, width, callback_steps, negative_prompt=None, prompt_embeds=None,
                     negative_prompt_embeds=None, ip_adapter_image=None, ip_adapter_image_embeds=None,
                     callback_on_step_end_tensor_inputs=None):
        """diffusers ``StableDiffusionPipeline.check_inputs`` as called at models.py:103-114: same conditions, same
        ``ValueError`` texts."""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is not None and (not isinstance(callback_steps, int) or callback_steps <= 0):
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")
        if callback_on_step_end_tensor_inputs is not None and not all(
                k in self._callback_tensor_inputs for k in callback_on_step_end_tensor_inputs):
            raise ValueError(
                f"`callback_on_step_end_tensor_inputs` has to be in {self._callback_tensor_inputs}, but found "
                f"{[k for k in callback_on_step_end_tensor_inputs if k not in self._callback_tensor_inputs]}")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `prompt`: {prompt} and `prompt_embeds`: {prompt_embeds}. Please make "
                             "sure to only forward one of the two.")
        elif prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` "
                             "undefined.")
        elif prompt is not None and (not isinstance(prompt, str) and not isinstance(prompt, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError(f"Cannot forward both `negative_prompt`: {negative_prompt} and `negative_prompt_embeds`:"
                             f" {negative_prompt_embeds}. Please make sure to only forward one of the two.")
        if prompt_embeds is not None and negative_prompt_embeds is not None:
            if prompt_embeds.shape != negative_prompt_embeds.shape:
                raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed "
                                 f"directly, but got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds`"
                                 f" {negative_prompt_embeds.shape}.")
        if ip_adapter_image is not None and ip_adapter_image_embeds is not None:
            raise ValueError("Provide either `ip_adapter_image` or `ip_adapter_image_embeds`. Cannot leave both "
                             "`ip_adapter_image` and `ip_adapter_image_embeds` defined.")

    def _check_call(self, prompt, height, width, negative_prompt, prompt_embeds, negative_prompt_embeds, latents,
                    timesteps, sigmas, num_images_per_prompt, guidance_rescale, output_type, kwargs):
        """The head of the reference ``call`` (models.py:64-114) -- default size, ``check_inputs`` -- followed by what
        this engine does not implement: every such argument raises instead of being ignored."""
        # the deprecated pair (models.py:65-80): called by the single-scheduler loops only (:275-282, :1407-1414; the
        # two-scheduler and interleaved bodies accept and never call it, :641-650 / :1074-1083)
        self._legacy_callback = (kwargs.pop("callback", None), kwargs.get("callback_steps"))
        callback_steps = kwargs.pop("callback_steps", None)
        tensor_inputs = kwargs.pop("callback_on_step_end_tensor_inputs", ["latents"])
        self._callback_inputs = list(tensor_inputs)
        ip_image, ip_embeds = kwargs.pop("ip_adapter_image", None), kwargs.pop("ip_adapter_image_embeds", None)
        if not height or not width:                             # models.py:85-100: both default together
            height = width = self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps, negative_prompt, prompt_embeds, negative_prompt_embeds,
                          ip_image, ip_embeds, tensor_inputs)
        if output_type not in ("latent", "pt", "np", "pil"):
            raise ValueError(f"output_type {output_type!r}: expected one of 'latent', 'pt', 'np', 'pil'")
        want = self.latent_size * self.vae_scale_factor
        if (height, width) != (want, want):
            raise NotImplementedError(f"height x width = {height} x {width}: this pipeline's engines are recorded for "
                                      f"{want} x {want} (latent_size={self.latent_size}); build it with another "
                                      "latent_size for other resolutions")
        if timesteps is not None and sigmas is not None:
            raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
        if sigmas is not None:
            raise ValueError(f"The current scheduler class {self.scheduler.__class__}'s `set_timesteps` does not support "
                             "custom sigmas schedules. Please check whether you are using the correct scheduler.")
        self._guidance_rescale = float(guidance_rescale or 0.0)   # models.py:117
        n_img = 1 if num_images_per_prompt is None else num_images_per_prompt
        if not isinstance(n_img, int) or isinstance(n_img, bool) or n_img < 1:
            raise ValueError(f"`num_images_per_prompt` has to be a positive integer but is {num_images_per_prompt!r}")
        self._images_per_prompt = n_img
        unsupported = {"ip_adapter_image": ip_image is not None,
                       "ip_adapter_image_embeds": ip_embeds is not None,
                       "cross_attention_kwargs": kwargs.pop("cross_attention_kwargs", None) is not None,
                       "clip_skip": kwargs.pop("clip_skip", None) is not None}
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"{', '.join(bad)}: not used by any reference driver and not implemented by the "
                                      "B200 engine")
        if latents is not None:
            batch = self._batch_size(prompt, prompt_embeds)
            shape = (batch, self.arch.in_channels, self.latent_size, self.latent_size)
            if tuple(latents.shape) != shape:
                raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")

    # ---------------------------------------------------------------- engine plumbing
    def engine(self, n_latents, cfg_dup) -> UNetEngine:
        if self.device.type != "cuda":
            raise RuntimeError("the B200 sampling engine needs a CUDA device: call .to('cuda') first "
                               "(there is no CPU fallback)")
        if self._weights is None:
            self._weights = PackedWeights(self._unet_sd, self.device)
        branch = self._deepcache["branch"] if self._deepcache else 0
        key = (n_latents, bool(cfg_dup), self.dtype, branch)
        if key not in self._engines:
            eng = UNetEngine(self._weights, n_latents=n_latents, cfg_dup=cfg_dup, arch=self.arch,
                             height=self.latent_size, width=self.latent_size, io_dtype=self.dtype,
                             device=self.device, cache_branch=branch)
            if self.use_cuda_graphs:
                eng.capture_graphs()
            self._engines[key] = eng
        return self._engines[key]

    def _encode(self, prompts):
        """CLIP text tower (models.py:139-149) on the native engine (clip_engine.ClipTextEngine); no library path."""
        if self.device.type != "cuda":
            raise RuntimeError("the prompt encoder runs on the B200 engine: call .to('cuda') first")
        from .clip_engine import ClipTextEngine

        key = ("text", len(prompts))
        if key not in self._engines:
            c = self.text_encoder.config
            self._engines[key] = ClipTextEngine(self.text_encoder.state_dict(), n=len(prompts),
                                                seq=c.max_position_embeddings, width=c.hidden_size,
                                                heads=c.num_attention_heads, layers=c.num_hidden_layers,
                                                mlp=c.intermediate_size, device=self.device)
        ids, _ = self.tokenizer(list(prompts))
        return self._engines[key].last_hidden_state(ids).clone()

    def _batch_size(self, prompt, prompt_embeds):
        """Latents per call: prompts x ``num_images_per_prompt`` (models.py:123-128 and ``batch_size *
        num_images_per_prompt`` at :173)."""
        prompts = (1 if isinstance(prompt, str) else len(prompt)) if prompt is not None else prompt_embeds.shape[0]
        return prompts * getattr(self, "_images_per_prompt", 1)

    def _context(self, pe, ne, do_cfg):
        """UNet context: ``cat([negative, positive])`` under guidance (models.py:154-155), every prompt's embedding
        repeated ``num_images_per_prompt`` times in place (diffusers ``encode_prompt``: ``repeat(1, n, 1).view(B * n,
        seq, -1)``, i.e. the n images of a prompt are adjacent)."""
        n = getattr(self, "_images_per_prompt", 1)
        if n > 1:
            pe = pe.repeat_interleave(n, dim=0)
            ne = ne.repeat_interleave(n, dim=0) if ne is not None else None
        # what ``callback_on_step_end`` may ask for by name (models.py:263-267): the loop's locals of those names
        self._callback_tensors = {"prompt_embeds": torch.cat([ne, pe]) if do_cfg else pe, "negative_prompt_embeds": ne}
        return self._callback_tensors["prompt_embeds"]

    def encode_prompt(self, prompt, do_cfg, prompt_embeds=None, negative_prompt_embeds=None, negative_prompt=None):
        if prompt_embeds is None:
            prompts = [prompt] if isinstance(prompt, str) else list(prompt)
            prompt_embeds = self._encode(prompts)
        prompt_embeds = prompt_embeds.to(device=self.device, dtype=torch.bfloat16)
        if do_cfg and negative_prompt_embeds is None:
            n = prompt_embeds.shape[0]
            neg = negative_prompt if negative_prompt is not None else ""
            neg = [neg] * n if isinstance(neg, str) else list(neg)
            negative_prompt_embeds = self._encode(neg)
        if negative_prompt_embeds is not None:
            negative_prompt_embeds = negative_prompt_embeds.to(device=self.device, dtype=torch.bfloat16)
        return prompt_embeds, negative_prompt_embeds

    def prepare_latents(self, batch, generator, latents, init_noise_sigma, rng_rows=None):
        """``prepare_latents`` / ``randn_tensor`` of models.py:172-182.  ``rng_rows = (lo, hi, total)``: this call
        computes rows lo:hi of a GLOBAL batch of ``total`` prompts -- the full tensor is drawn (what a single
        process would draw) and sliced, so a sharded run reproduces the single-GPU noise."""
        shape = (batch, self.arch.in_channels, self.latent_size, self.latent_size)
        if latents is None:
            if rng_rows is not None:
                lo, hi, total = rng_rows
                assert hi - lo == batch, (rng_rows, batch)
                latents = S.randn_tensor((total,) + shape[1:], generator=generator, device=self.device,
                                         dtype=self.dtype)[lo:hi]
            else:
                latents = S.randn_tensor(shape, generator=generator, device=self.device, dtype=self.dtype)
        else:
            latents = latents.to(device=self.device, dtype=self.dtype)
        return latents * init_noise_sigma

    def _replay_rng(self, plan, batch, latents, generator, init_noise_sigma):
        """``rng_only=True``: consume from ``generator`` exactly what the call would -- the initial latents and
        every noisy step's draw, in order -- with no text encoding, no engine and no kernel launch.  This is how
        a rank walks over the batches other ranks compute (experiments/base_experiment.py ``generate``)."""
        if self.device.type != "cuda" and generator is not None and generator.device.type == "cuda":
            raise RuntimeError("replaying a CUDA generator needs the pipeline on that device: call .to('cuda')")
        self.prepare_latents(batch, generator, latents, init_noise_sigma)
        shape = (batch, self.arch.in_channels, self.latent_size, self.latent_size)
        for sched, t, extra in plan:
            sched.replay(t, shape, self.dtype, self.device, **extra)
        return None, 0.0, []

    @staticmethod
    def _extra_step_kwargs(scheduler, generator, eta):
        params = inspect.signature(scheduler._step).parameters
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    def _is_cached_step(self, t_list, i):
        """DeepCache schedule (appendix A.4): cur = first index of t in the timestep list."""
        dc = self._deepcache
        if not dc:
            return False
        cur = t_list.index(t_list[i])
        if dc.get("start") is None:
            dc["start"] = cur
        return (cur - dc["start"]) % dc["interval"] != 0

    def _guided(self, eps, B, do_cfg, guidance_scale):
        """What the fused update is given: ``(eps_uncond, eps_text, g)`` -- the kernel forms ``u + g (c - u)`` itself
        (models.py:238-242) -- or ``(eps, None, 0)`` without guidance.  ``guidance_rescale`` > 0 (models.py:244-250,
        diffusers ``rescale_noise_cfg``; 0 in every reference driver) needs two per-image standard deviations between
        the combine and the update: that branch forms the guided prediction with the same torch expressions as the
        reference, on the device, and hands the update kernel the finished prediction."""
        if not do_cfg:
            return eps, None, 0.0
        u, c = eps[:B], eps[B:]
        phi = self.guidance_rescale
        if phi <= 0.0:
            return u, c, guidance_scale
        uf, cf = u.float(), c.float()
        cfg = uf + guidance_scale * (cf - uf)
        dims = list(range(1, cfg.ndim))
        rescaled = cfg * (cf.std(dim=dims, keepdim=True) / cfg.std(dim=dims, keepdim=True))
        return (phi * rescaled + (1 - phi) * cfg).to(eps.dtype), None, 0.0

    def _denoise_step(self, eng, scheduler, t, do_cfg, guidance_scale, cached, extra):
        """models.py:217-261 for one timestep: UNet plan replay + one fused update kernel."""
        eps = eng.forward(float(t), cached=cached)
        eps_u, eps_c, g = self._guided(eps, eng.n_lat, do_cfg, guidance_scale)
        if eps_c is not None:
            return scheduler.step_cfg(eps_u, eps_c, g, t, eng.x_in, out=eng.x_in, **extra)
        return scheduler._step(eps_u, None, 0.0, t, eng.x_in, out=eng.x_in, **extra)

    def _x0_mode(self, schedulers, output_type):
        """The loop keeps ``x0_pred[0]`` only (models.py:257-261) and ``output_type="latent"`` never decodes it:
        tell the fused step to write one image's x0, or none."""
        for s_ in schedulers:
            s_.x0_rows, s_.skip_x0 = 1, (output_type == "latent" or not self.decode_x0_preds)

    @staticmethod
    def _x0_reset(schedulers):
        for s_ in schedulers:
            s_.x0_rows, s_.skip_x0, s_.rng_rows = None, False, None

    def _run_callback(self, callback, eng, i, t):
        """``callback_on_step_end(pipe, i, t, {name: tensor for name in callback_on_step_end_tensor_inputs})`` as at
        models.py:263-273.  A returned ``latents`` replaces the resident latents (teacher-forced parity runs use it), a
        returned ``prompt_embeds`` -- the UNet context, ``[negative, positive]`` under guidance -- replaces the
        context of every later step; a returned ``negative_prompt_embeds`` is kept for later callbacks and, as in the
        reference, feeds nothing else."""
        if callback is None:
            return
        held = self._callback_tensors
        names = getattr(self, "_callback_inputs", ["latents"])
        out = callback(self, i, t, {k: eng.x_in if k == "latents" else held[k] for k in names}) or {}
        new = out.get("latents")
        if new is not None and new.data_ptr() != eng.x_in.data_ptr():
            eng.x_in.copy_(new)
        ctx = out.get("prompt_embeds")
        if ctx is not None and ctx is not held["prompt_embeds"]:
            held["prompt_embeds"] = ctx.to(device=held["prompt_embeds"].device, dtype=held["prompt_embeds"].dtype)
            eng.set_context(held["prompt_embeds"])
        if out.get("negative_prompt_embeds") is not None:
            held["negative_prompt_embeds"] = out["negative_prompt_embeds"]

    def _run_legacy_callback(self, i, t, eng, n_timesteps, num_inference_steps):
        """The deprecated ``callback(step_idx, t, latents)`` every ``callback_steps`` indices, under the progress-bar
        condition of models.py:275-282 (``num_warmup_steps`` of :205: PLMS' grid has one timestep more than steps)."""
        callback, callback_steps = getattr(self, "_legacy_callback", (None, None))
        if callback is None:
            return
        order = getattr(self.scheduler, "order", 1)
        num_warmup_steps = n_timesteps - num_inference_steps * order
        if i == n_timesteps - 1 or ((i + 1) > num_warmup_steps and (i + 1) % order == 0):
            if i % callback_steps == 0:
                callback(i // order, t, eng.x_in)

    def vae_engine(self, n_img):
        """Native decoder plan for ``n_img`` latents (vae_engine.VaeEngine), built on first use."""
        from .vae_engine import VaeEngine

        key = ("vae", n_img, self.dtype)
        if key not in self._engines:
            self._engines[key] = VaeEngine(dict(self.vae.state_dict()), n_img=n_img,
                                           latent=self.latent_size, io_dtype=self.dtype, device=self.device)
        return self._engines[key]

    def _decode(self, z):
        """``vae.decode`` (models.py:288-302) on the native engine (vae_engine.VaeEngine); no library path."""
        if self.device.type != "cuda":
            raise RuntimeError("the VAE decoder runs on the B200 engine: call .to('cuda') first")
        return self.vae_engine(z.shape[0]).decode(z.to(self.dtype)).clone()

    def _finish(self, eng, x0_preds, output_type, exec_time, return_dict=True):
        """models.py:287-335: decode, ``image_processor.postprocess`` (denormalise + output type), return arity."""
        latents = eng.x_in.clone()
        images_x0 = []
        if output_type == "latent":
            image = latents
        else:
            sf = self.vae.config.scaling_factor
            image = postprocess_images((self._decode(latents / sf) / 2 + 0.5).clamp(0, 1), output_type)
            if self.decode_x0_preds:
                for x0 in x0_preds:
                    images_x0.append(postprocess_images((self._decode(x0 / sf) / 2 + 0.5).clamp(0, 1), output_type))
        if not return_dict:
            return (image, None), exec_time, images_x0
        return PipelineOutput(images=image, nsfw_content_detected=None), exec_time, images_x0

    def __call__(self, *args, return_execution_time=True, **kwargs):
        result, execution_time, x0_preds = self.call(*args, **kwargs)
        if return_execution_time:
            return result, execution_time, x0_preds
        return result, x0_preds


@models_registry.add_to_registry("stable_diffusion_model")
class StableDiffusionModel(_PipelineBase):
    @torch.no_grad()
    def call(self, prompt=None, height=None, width=None, num_inference_steps: int = 50, timesteps=None,
             sigmas=None, guidance_scale: float = 7.5, negative_prompt=None, num_images_per_prompt=1,
             eta: float = 0.0, generator=None, latents=None, prompt_embeds=None, negative_prompt_embeds=None,
             output_type="pil", return_dict=True, guidance_rescale: float = 0.0, skip_timesteps=(),
             callback_on_step_end=None, **kwargs):
        rng_only, rng_rows = kwargs.pop("rng_only", False), kwargs.pop("rng_rows", None)
        self._check_call(prompt, height, width, negative_prompt, prompt_embeds, negative_prompt_embeds, latents,
                         timesteps, sigmas, num_images_per_prompt, guidance_rescale, output_type, kwargs)
        self._guidance_scale = guidance_scale
        batch = self._batch_size(prompt, prompt_embeds)
        do_cfg = self.do_classifier_free_guidance
        ts, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, self.device, timesteps)
        t_list = [int(t) for t in ts.tolist()]
        skip = set(skip_timesteps or ())            # loop INDICES, models.py:1220-1223,1338-1340
        extra = self._extra_step_kwargs(self.scheduler, generator, eta)
        self.scheduler.rng_rows = rng_rows
        if rng_only:
            plan = [(self.scheduler, t, extra) for i, t in enumerate(t_list) if i not in skip]
            return self._replay_rng(plan, batch, latents, generator, self.scheduler.init_noise_sigma)
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = self._context(pe, ne, do_cfg)
        latents = self.prepare_latents(batch, generator, latents, self.scheduler.init_noise_sigma, rng_rows)
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        if self._deepcache:
            self._deepcache["start"] = None
        self._num_timesteps = len(t_list)
        self.last_step_kinds = []
        x0_preds = []
        self._x0_mode([self.scheduler], output_type)
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(t_list):
            if i in skip:                           # skip-steps variant, models.py:1338-1340
                continue
            cached = self._is_cached_step(t_list, i)
            self.last_step_kinds.append("cached" if cached else "full")
            step = self._denoise_step(eng, self.scheduler, t, do_cfg, guidance_scale, cached, extra)
            if len(step) == 2 and step[1] is not None:
                x0_preds.append(step[1][0:1])
            self._run_callback(callback_on_step_end, eng, i, t)
            self._run_legacy_callback(i, t, eng, len(t_list), num_inference_steps)
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        self._x0_reset([self.scheduler])
        return self._finish(eng, x0_preds, output_type, exec_time, return_dict)


@models_registry.add_to_registry("stable_diffusion_model_skip_timesteps")
class StableDiffusionModelSkipTimesteps(StableDiffusionModel):
    """models.py:1138-1467: the single-scheduler loop with ``if i in skip_timesteps: continue``."""

    def call(self, *args, skip_timesteps=None, **kwargs):
        # models.py:1220-1223: ``None`` -> nothing skipped; entries are loop indices, not timestep values
        return super().call(*args, skip_timesteps=tuple(skip_timesteps or ()), **kwargs)


@models_registry.add_to_registry("stable_diffusion_model_two_schedulers")
class StableDiffusionModelTwoSchedulers(_PipelineBase):
    scheduler_first = None
    scheduler_second = None

    @torch.no_grad()
    def call(self, prompt=None, height=None, width=None, num_inference_steps_first: int = 50,
             num_inference_steps_second: int = 50, num_step_switch: int = 10, type_switch: str = "closest",
             timesteps=None, sigmas=None, guidance_scale: float = 7.5, negative_prompt=None,
             num_images_per_prompt=1, eta: float = 0.0, generator=None, latents=None, prompt_embeds=None,
             negative_prompt_embeds=None, output_type="pil", return_dict=True, guidance_rescale: float = 0.0,
             callback_on_step_end=None, **kwargs):
        rng_only, rng_rows = kwargs.pop("rng_only", False), kwargs.pop("rng_rows", None)
        self._check_call(prompt, height, width, negative_prompt, prompt_embeds, negative_prompt_embeds, latents,
                         timesteps, sigmas, num_images_per_prompt, guidance_rescale, output_type, kwargs)
        self._guidance_scale = guidance_scale
        batch = self._batch_size(prompt, prompt_embeds)
        do_cfg = self.do_classifier_free_guidance
        # models.py:487-494: the second scheduler runs on the FIRST scheduler's grid (N2 unused)
        ts1, _ = retrieve_timesteps(self.scheduler_first, num_inference_steps_first, self.device, timesteps)
        ts2, _ = retrieve_timesteps(self.scheduler_second, device=self.device, timesteps=ts1.cpu().numpy())
        first, second = self.switch_timestamp(ts1, ts2, num_step_switch, type_switch)
        # models.py:520: ONE extra_step_kwargs, derived from ``self.scheduler`` (the pipeline's default), serves
        # both phases; a keyword the phase's scheduler does not take is dropped instead of raising TypeError
        extra = self._extra_step_kwargs(self.scheduler, generator, eta)
        extra1 = {k: v for k, v in extra.items() if k in inspect.signature(self.scheduler_first._step).parameters}
        extra2 = {k: v for k, v in extra.items() if k in inspect.signature(self.scheduler_second._step).parameters}
        self.scheduler_first.rng_rows = self.scheduler_second.rng_rows = rng_rows
        if rng_only:
            plan = [(self.scheduler_first, int(t), extra1) for t in first] + \
                   [(self.scheduler_second, int(t), extra2) for t in second]
            return self._replay_rng(plan, batch, latents, generator, self.scheduler_first.init_noise_sigma)
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = self._context(pe, ne, do_cfg)
        latents = self.prepare_latents(batch, generator, latents, self.scheduler_first.init_noise_sigma, rng_rows)
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        self._num_timesteps = len(first) + len(second)
        self.last_timesteps = (list(first), list(second))
        x0_preds = []
        self._x0_mode([self.scheduler_first, self.scheduler_second], output_type)
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(first + second):     # models.py:545-621 (history seeding :603-611 is a
            in_first = i < len(first)              # no-op for solver_order <= 2, SURVEY C-4)
            sched, extra = (self.scheduler_first, extra1) if in_first else (self.scheduler_second, extra2)
            step = self._denoise_step(eng, sched, int(t), do_cfg, guidance_scale, False, extra)
            if len(step) == 2 and step[1] is not None:
                x0_preds.append(step[1][0:1])
            self._run_callback(callback_on_step_end, eng, i, int(t))
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        self._x0_reset([self.scheduler_first, self.scheduler_second])
        return self._finish(eng, x0_preds, output_type, exec_time, return_dict)

    def switch_timestamp(self, timesteps_first, timesteps_second, num_step_switch, type_switch="closest"):
        """models.py:704-730 -> python lists of np.int64 (bit-exact integer schedule)."""
        first = list(timesteps_first[:num_step_switch].cpu().numpy())
        second_all = timesteps_second.cpu().numpy()
        pivot = first[-1]
        if type_switch == "closest":
            k = int(np.argmin([abs(int(t) - int(pivot)) for t in second_all]))
        elif type_switch == "left_closest":
            k = [i for i, t in enumerate(second_all) if int(t) - int(pivot) >= 0][-1]
        elif type_switch == "right_closest":
            k = [i for i, t in enumerate(second_all) if int(t) - int(pivot) <= 0][0]
        else:
            raise ValueError(f"unknown type_switch {type_switch!r}")
        return first, list(second_all[k:])


@models_registry.add_to_registry("stable_diffusion_model_interliving_schedulers")
class StableDiffusionModelInterlivingSchedulers(_PipelineBase):
    """src/models.py:733-1135: a multistep main scheduler whose grid is grouped ``solver_order`` steps at a time;
    the groups listed in ``interliving_steps`` are replaced by ONE step of the inter scheduler (set up on
    N // order steps, so e.g. a DDIM step spans the whole group), and after every step the other scheduler's
    multistep history is fed the converted model output (``feed_history``)."""

    scheduler_main = None
    scheduler_inter = None

    @staticmethod
    def partition(timesteps_main, solver_order, interliving_steps):
        """models.py:944-961 -> (timesteps that are evaluated, those of them the inter scheduler takes)."""
        kept, inter = [], []
        for i, t in enumerate(int(v) for v in timesteps_main):
            if i // solver_order in interliving_steps:
                if i % solver_order != 0:
                    continue
                inter.append(t)
            kept.append(t)
        return kept, inter

    @torch.no_grad()
    def call(self, prompt=None, height=None, width=None, num_inference_steps: int = 50, interliving_steps=None,
             timesteps=None, sigmas=None, guidance_scale: float = 7.5, negative_prompt=None,
             num_images_per_prompt=1, eta: float = 0.0, generator=None, latents=None, prompt_embeds=None,
             negative_prompt_embeds=None, output_type="pil", return_dict=True, guidance_rescale: float = 0.0,
             callback_on_step_end=None, **kwargs):
        main, inter_s = self.scheduler_main, self.scheduler_inter
        if main is None or inter_s is None:
            raise ValueError("scheduler_main / scheduler_inter must be set (interliving_exp.py:40-62)")
        interliving_steps = list(interliving_steps or [])
        rng_only, rng_rows = kwargs.pop("rng_only", False), kwargs.pop("rng_rows", None)
        self._check_call(prompt, height, width, negative_prompt, prompt_embeds, negative_prompt_embeds, latents,
                         timesteps, sigmas, num_images_per_prompt, guidance_rescale, output_type, kwargs)
        self._guidance_scale = guidance_scale
        batch = self._batch_size(prompt, prompt_embeds)
        do_cfg = self.do_classifier_free_guidance
        order = main.config.solver_order
        ts_main, _ = retrieve_timesteps(main, num_inference_steps, self.device, timesteps)        # models.py:880-886
        retrieve_timesteps(inter_s, num_inference_steps // order, self.device, timesteps)         # models.py:888-894
        kept, t_inter = self.partition(ts_main.tolist(), order, interliving_steps)
        self._num_timesteps = len(ts_main) - len(interliving_steps)                               # models.py:939
        self.last_timesteps = (kept, t_inter)
        extra = self._extra_step_kwargs(self.scheduler, generator, eta)                           # models.py:911
        extra_main = {k: v for k, v in extra.items() if k in inspect.signature(main._step).parameters}
        extra_inter = {k: v for k, v in extra.items() if k in inspect.signature(inter_s._step).parameters}
        main.rng_rows = inter_s.rng_rows = rng_rows
        if rng_only:
            plan = [(inter_s, t, extra_inter) if t in t_inter else (main, t, extra_main) for t in kept]
            return self._replay_rng(plan, batch, latents, generator, self.scheduler.init_noise_sigma)
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = self._context(pe, ne, do_cfg)
        latents = self.prepare_latents(batch, generator, latents, self.scheduler.init_noise_sigma, rng_rows)
        feed_inter = isinstance(inter_s, S.DPMSolverScheduler)                                    # models.py:1045
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        B = eng.n_lat
        x0_preds = []
        self._x0_mode([main, inter_s], output_type)
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(kept):
            stepper, extra, other = (inter_s, extra_inter, main) if t in t_inter else (main, extra_main,
                                                                                       inter_s if feed_inter else None)
            eps = eng.forward(float(t))
            eps_u, eps_c, g = self._guided(eps, B, do_cfg, guidance_scale)
            step = stepper._step(eps_u, eps_c, g, t, eng.x_in, out=eng.x_in, **extra)
            if len(step) == 2 and step[1] is not None:
                x0_preds.append(step[1][0:1])
            if other is not None:
                other.feed_history(eps_u, eps_c, g, eng.x_in)       # post-step latents, pre-step noise: as written
            self._run_callback(callback_on_step_end, eng, i, t)
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        self._x0_reset([main, inter_s])
        return self._finish(eng, x0_preds, output_type, exec_time, return_dict)
