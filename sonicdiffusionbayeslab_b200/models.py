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
import os
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import schedulers as S
from .registry import models_registry
from .text import encode_prompts, load_tokenizer, make_text_encoder
from .unet_engine import PackedWeights, UNetArch, UNetEngine
from .unet_spec import random_unet_state_dict, validate_state_dict
from .vae import make_vae

# runwayml/stable-diffusion-v1-5 scheduler/scheduler_config.json + instantiated defaults
SD15_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    trained_betas=None, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1, clip_sample=False,
    prediction_type="epsilon", timestep_spacing="leading",
)


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
        self.use_native_vae = True
        self.use_native_text = True
        self.decode_x0_preds = True
        self.unet = SimpleNamespace(config=SimpleNamespace(
            in_channels=arch.in_channels, sample_size=latent_size, time_cond_proj_dim=None))
        self.last_step_kinds = []                   # "full"/"cached" per UNet call of the last run

    # ---------------------------------------------------------------- construction
    @classmethod
    def from_pretrained(cls, pretrained_model, timestamps=None, safety_checker=None,
                        requires_safety_checker=False, torch_dtype=torch.bfloat16, seed=29, **kwargs):
        """``pretrained_model``: a local diffusers-layout directory (unet/, vae/, text_encoder/,
        tokenizer/) is loaded; any other id (e.g. "runwayml/stable-diffusion-v1-5" without network)
        yields seeded random-init weights of the SD-v1.5 architecture."""
        if torch_dtype == torch.float16:
            torch_dtype = torch.bfloat16            # the engine computes in bf16 (SURVEY C-11)
        arch = UNetArch()
        root = pretrained_model if isinstance(pretrained_model, str) and os.path.isdir(pretrained_model) else None
        sub = (lambda *p: os.path.join(root, *p)) if root else None
        unet_file = sub("unet", "diffusion_pytorch_model.safetensors") if root else None
        if unet_file and os.path.exists(unet_file):
            sd = _load_safetensors(unet_file)
            validate_state_dict(sd, arch)
        else:
            sd = random_unet_state_dict(seed, arch)
        vae = make_vae(seed, dtype=torch_dtype)
        vae_file = sub("vae", "diffusion_pytorch_model.safetensors") if root else None
        if vae_file and os.path.exists(vae_file):
            vae.load_diffusers_state_dict(_load_safetensors(vae_file))
        text = make_text_encoder(seed, sub("text_encoder") if root else None, dtype=torch_dtype)
        tok = load_tokenizer(sub("tokenizer") if root else None)
        sched = S.PNDMScheduler.from_config(SD15_SCHEDULER_CONFIG)
        return cls(sd, vae, text, tok, sched, arch=arch, torch_dtype=torch_dtype, timestamps=timestamps, seed=seed)

    def to(self, device):
        """Weights are packed into HBM on the first move to a CUDA device and stay resident; moving
        "to cpu" (the reference bounces the model around every sweep point, ddim.py:31-33) keeps
        the engines alive and is a no-op."""
        dev = torch.device(device)
        if dev.type == "cuda":
            self.device = dev
            self.vae.to(dev)
            self.text_encoder.to(dev)
        return self

    @property
    def num_timesteps(self):
        return self._num_timesteps

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 1 and self.unet.config.time_cond_proj_dim is None

    def load_lora_weights(self, adapter, adapter_scale=1.0, **_):
        """Offline LoRA merge source (consistency_model.py:20-21): a local .safetensors with
        ``<key>.lora_A/.lora_B(.alpha)`` tensors; unknown ids (no network) leave weights unchanged."""
        self._lora = None
        if isinstance(adapter, str) and os.path.exists(adapter):
            self._lora = (_load_safetensors(adapter), adapter_scale)

    def fuse_lora(self, lora_scale=1.0):
        lora = getattr(self, "_lora", None)
        if not lora:
            return
        tensors, scale = lora
        for k in [k for k in tensors if k.endswith("lora_A.weight")]:
            base = k[: -len(".lora_A.weight")]
            target = base.replace("unet.", "", 1) + ".weight"
            if target not in self._unet_sd:
                continue
            A, B = tensors[k].float(), tensors[base + ".lora_B.weight"].float()
            alpha = float(tensors.get(base + ".alpha", torch.tensor(float(A.shape[0]))))
            delta = (B.flatten(1) @ A.flatten(1)) * (alpha / A.shape[0]) * scale * lora_scale
            self._unet_sd[target] = self._unet_sd[target].float() + delta.reshape(self._unet_sd[target].shape)
        self._weights = None
        self._engines = {}

    # ---------------------------------------------------------------- engine plumbing
    def engine(self, n_latents, cfg_dup) -> UNetEngine:
        if self.device.type != "cuda":
            raise RuntimeError("the B200 sampling engine needs a CUDA device: call .to('cuda') first "
                               "(there is no CPU fallback)")
        if self._weights is None:
            self._weights = PackedWeights(self._unet_sd, self.device)
        key = (n_latents, bool(cfg_dup), self.dtype)
        if key not in self._engines:
            eng = UNetEngine(self._weights, n_latents=n_latents, cfg_dup=cfg_dup, arch=self.arch,
                             height=self.latent_size, width=self.latent_size, io_dtype=self.dtype,
                             device=self.device)
            if self.use_cuda_graphs:
                eng.capture_graphs()
            self._engines[key] = eng
        return self._engines[key]

    def _encode(self, prompts):
        """CLIP text tower (models.py:139-149) on the native engine (clip_engine.ClipTextEngine)."""
        if not (self.use_native_text and self.device.type == "cuda"):
            return encode_prompts(self.tokenizer, self.text_encoder, prompts, self.device)
        from .clip_engine import ClipTextEngine

        key = ("text", len(prompts))
        if key not in self._engines:
            c = self.text_encoder.config
            sd = {k: v.detach() for k, v in self.text_encoder.state_dict().items()}
            self._engines[key] = ClipTextEngine(sd, n=len(prompts), seq=c.max_position_embeddings, width=c.hidden_size,
                                                heads=c.num_attention_heads, layers=c.num_hidden_layers,
                                                mlp=c.intermediate_size, device=self.device)
        ids, _ = self.tokenizer(list(prompts))
        return self._engines[key].last_hidden_state(ids).clone()

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

    def prepare_latents(self, batch, generator, latents, init_noise_sigma):
        shape = (batch, self.arch.in_channels, self.latent_size, self.latent_size)
        if latents is None:
            latents = S.randn_tensor(shape, generator=generator, device=self.device, dtype=self.dtype)
        else:
            latents = latents.to(device=self.device, dtype=self.dtype)
        return latents * init_noise_sigma

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

    def _denoise_step(self, eng, scheduler, t, do_cfg, guidance_scale, cached, extra):
        """models.py:217-261 for one timestep: UNet plan replay + one fused update kernel."""
        eps = eng.forward(float(t), cached=cached)
        B = eng.n_lat
        if do_cfg:
            return scheduler.step_cfg(eps[:B], eps[B:], guidance_scale, t, eng.x_in, out=eng.x_in, **extra)
        return scheduler._step(eps, None, 0.0, t, eng.x_in, out=eng.x_in, **extra)

    def _run_callback(self, callback, eng, i, t):
        """``callback_on_step_end(pipe, i, t, {"latents": ...})`` as at models.py:263-273; a returned
        ``{"latents": tensor}`` replaces the resident latents (used for teacher-forced parity runs)."""
        if callback is None:
            return
        out = callback(self, i, t, {"latents": eng.x_in}) or {}
        new = out.get("latents")
        if new is not None and new.data_ptr() != eng.x_in.data_ptr():
            eng.x_in.copy_(new)

    def vae_engine(self, n_img):
        """Native decoder plan for ``n_img`` latents (vae_engine.VaeEngine), built on first use."""
        from .vae_engine import VaeEngine

        key = ("vae", n_img, self.dtype)
        if key not in self._engines:
            self._engines[key] = VaeEngine({k: v.detach() for k, v in self.vae.state_dict().items()}, n_img=n_img,
                                           latent=self.latent_size, io_dtype=self.dtype, device=self.device)
        return self._engines[key]

    def _decode(self, z):
        """``vae.decode`` (models.py:288-302) on the native engine; ``use_native_vae = False`` keeps the
        PyTorch module (library kernels)."""
        if self.use_native_vae and self.device.type == "cuda":
            return self.vae_engine(z.shape[0]).decode(z.to(self.dtype)).clone()
        return self.vae.decode(z)[0]

    def _finish(self, eng, x0_preds, output_type, exec_time):
        latents = eng.x_in.clone()
        images_x0 = []
        if output_type == "latent":
            image = latents
        else:
            sf = self.vae.config.scaling_factor
            image = self._decode(latents / sf)
            image = (image / 2 + 0.5).clamp(0, 1)
            if self.decode_x0_preds:
                for x0 in x0_preds:
                    images_x0.append((self._decode(x0 / sf) / 2 + 0.5).clamp(0, 1))
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
        if guidance_rescale:
            raise NotImplementedError("guidance_rescale is 0 in every reference config and is not fused")
        if output_type not in ("pt", "latent"):
            raise NotImplementedError("output_type must be 'pt' (as the experiments use) or 'latent'")
        self._guidance_scale = guidance_scale
        batch = (1 if isinstance(prompt, str) else len(prompt)) if prompt is not None else prompt_embeds.shape[0]
        do_cfg = self.do_classifier_free_guidance
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = torch.cat([ne, pe]) if do_cfg else pe
        ts, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, self.device, timesteps)
        t_list = [int(t) for t in ts.tolist()]
        latents = self.prepare_latents(batch, generator, latents, self.scheduler.init_noise_sigma)
        extra = self._extra_step_kwargs(self.scheduler, generator, eta)
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        if self._deepcache:
            self._deepcache["start"] = None
        self._num_timesteps = len(t_list)
        self.last_step_kinds = []
        x0_preds = []
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(t_list):
            if i in skip_timesteps:                 # skip-steps variant, models.py:1338-1340
                continue
            cached = self._is_cached_step(t_list, i)
            self.last_step_kinds.append("cached" if cached else "full")
            step = self._denoise_step(eng, self.scheduler, t, do_cfg, guidance_scale, cached, extra)
            if len(step) == 2:
                x0_preds.append(step[1][0:1])
            self._run_callback(callback_on_step_end, eng, i, t)
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        return self._finish(eng, x0_preds, output_type, exec_time)


@models_registry.add_to_registry("stable_diffusion_model_skip_timesteps")
class StableDiffusionModelSkipTimesteps(StableDiffusionModel):
    """models.py:1138-1467: the single-scheduler loop with ``if i in skip_timesteps: continue``."""

    def call(self, *args, skip_timesteps=None, **kwargs):
        skip = self.timestamps if skip_timesteps is None else skip_timesteps
        return super().call(*args, skip_timesteps=tuple(skip or ()), **kwargs)


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
        if output_type not in ("pt", "latent"):
            raise NotImplementedError("output_type must be 'pt' or 'latent'")
        self._guidance_scale = guidance_scale
        batch = (1 if isinstance(prompt, str) else len(prompt)) if prompt is not None else prompt_embeds.shape[0]
        do_cfg = self.do_classifier_free_guidance
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = torch.cat([ne, pe]) if do_cfg else pe
        # models.py:487-494: the second scheduler runs on the FIRST scheduler's grid (N2 unused)
        ts1, _ = retrieve_timesteps(self.scheduler_first, num_inference_steps_first, self.device, timesteps)
        ts2, _ = retrieve_timesteps(self.scheduler_second, device=self.device, timesteps=ts1.cpu().numpy())
        first, second = self.switch_timestamp(ts1, ts2, num_step_switch, type_switch)
        latents = self.prepare_latents(batch, generator, latents, self.scheduler_first.init_noise_sigma)
        extra1 = self._extra_step_kwargs(self.scheduler_first, generator, eta)
        extra2 = self._extra_step_kwargs(self.scheduler_second, generator, eta)
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        self._num_timesteps = len(first) + len(second)
        self.last_timesteps = (list(first), list(second))
        x0_preds = []
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(first + second):     # models.py:545-621 (history seeding :603-611 is a
            in_first = i < len(first)              # no-op for solver_order <= 2, SURVEY C-4)
            sched, extra = (self.scheduler_first, extra1) if in_first else (self.scheduler_second, extra2)
            step = self._denoise_step(eng, sched, int(t), do_cfg, guidance_scale, False, extra)
            if len(step) == 2:
                x0_preds.append(step[1][0:1])
            self._run_callback(callback_on_step_end, eng, i, int(t))
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        return self._finish(eng, x0_preds, output_type, exec_time)

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
        if output_type not in ("pt", "latent"):
            raise NotImplementedError("output_type must be 'pt' or 'latent'")
        main, inter_s = self.scheduler_main, self.scheduler_inter
        if main is None or inter_s is None:
            raise ValueError("scheduler_main / scheduler_inter must be set (interliving_exp.py:40-62)")
        interliving_steps = list(interliving_steps or [])
        self._guidance_scale = guidance_scale
        batch = (1 if isinstance(prompt, str) else len(prompt)) if prompt is not None else prompt_embeds.shape[0]
        do_cfg = self.do_classifier_free_guidance
        pe, ne = self.encode_prompt(prompt, do_cfg, prompt_embeds, negative_prompt_embeds, negative_prompt)
        ctx = torch.cat([ne, pe]) if do_cfg else pe
        order = main.config.solver_order
        ts_main, _ = retrieve_timesteps(main, num_inference_steps, self.device, timesteps)        # models.py:880-886
        retrieve_timesteps(inter_s, num_inference_steps // order, self.device, timesteps)         # models.py:888-894
        kept, t_inter = self.partition(ts_main.tolist(), order, interliving_steps)
        self._num_timesteps = len(ts_main) - len(interliving_steps)                               # models.py:939
        self.last_timesteps = (kept, t_inter)
        latents = self.prepare_latents(batch, generator, latents, self.scheduler.init_noise_sigma)
        extra_main = self._extra_step_kwargs(main, generator, eta)
        extra_inter = self._extra_step_kwargs(inter_s, generator, eta)
        feed_inter = isinstance(inter_s, S.DPMSolverScheduler)                                    # models.py:1045
        eng = self.engine(batch, do_cfg)
        eng.set_context(ctx)
        eng.x_in.copy_(latents)
        B = eng.n_lat
        x0_preds = []
        torch.cuda.synchronize(self.device)
        start = time.perf_counter()
        for i, t in enumerate(kept):
            stepper, extra, other = (inter_s, extra_inter, main) if t in t_inter else (main, extra_main,
                                                                                       inter_s if feed_inter else None)
            eps = eng.forward(float(t))
            eps_u, eps_c, g = (eps[:B], eps[B:], guidance_scale) if do_cfg else (eps, None, 0.0)
            step = stepper._step(eps_u, eps_c, g, t, eng.x_in, out=eng.x_in, **extra)
            if len(step) == 2:
                x0_preds.append(step[1][0:1])
            if other is not None:
                other.feed_history(eps_u, eps_c, g, eng.x_in)       # post-step latents, pre-step noise: as written
            self._run_callback(callback_on_step_end, eng, i, t)
        torch.cuda.synchronize(self.device)
        exec_time = time.perf_counter() - start
        return self._finish(eng, x0_preds, output_type, exec_time)
