"""Oracle restatement of the reference's denoising loops (test infrastructure only).

``denoise``      <- /root/reference/src/models.py:154-155,167-182,210-282 (single scheduler)
``denoise_two``  <- /root/reference/src/models.py:487-502,545-621,704-730 (two-scheduler switch)
``denoise_interleaved`` <- /root/reference/src/models.py:880-897,939-961,963-1053 (interleaved schedulers)
Same op order and dtypes as the reference: latents duplicated with ``torch.cat``, UNet call,
``uncond + g * (text - uncond)`` in the model dtype, ``scheduler.step``.  Optional DeepCache via
``oracle.deepcache.DeepCacheOracle`` and teacher forcing (``forced_latents``) for parity tests.
Pinned against the reference's own source where it has one (oracle/refexec.py; see oracle/__init__.py).
"""
from __future__ import annotations

import inspect

import numpy as np
import torch

from .schedulers import randn_tensor


def prepare_latents(shape, generator, device, dtype, init_noise_sigma=1.0, latents=None):
    if latents is None:
        latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
    else:
        latents = latents.to(device)
    return latents * init_noise_sigma


def _extra(scheduler, generator, eta):
    params = set(inspect.signature(scheduler.step).parameters)
    kw = {}
    if "eta" in params:
        kw["eta"] = eta
    if "generator" in params:
        kw["generator"] = generator
    return kw


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """diffusers ``rescale_noise_cfg`` as called at /root/reference/src/models.py:244-250 (and :583-589, :1001-1007,
    :1374-1380): the guided prediction is brought to the per-image standard deviation of the text prediction
    (``torch.std`` over all non-batch dims, Bessel-corrected) and blended with weight ``guidance_rescale``."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=dims, keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


def _edited(context_edit, i, ctx):
    """``callback_on_step_end`` may return new ``prompt_embeds`` (models.py:263-273): they replace the UNet context --
    under guidance the concatenated ``[negative, positive]`` batch, models.py:154-155 -- of every later step.
    ``context_edit(i, ctx)`` returns the replacement after loop index ``i``, or None."""
    if context_edit is None:
        return ctx
    new = context_edit(i, ctx)
    return ctx if new is None else new


def _guide(noise_pred, guidance_scale, guidance_rescale):
    """models.py:238-250: ``u + g (c - u)``, then the optional std rescale."""
    u, c = noise_pred.chunk(2)
    out = u + guidance_scale * (c - u)
    if guidance_rescale > 0.0:
        out = rescale_noise_cfg(out, c, guidance_rescale=guidance_rescale)
    return out


@torch.no_grad()
def denoise(unet, scheduler, prompt_embeds, negative_prompt_embeds, latents, num_inference_steps,
            guidance_scale=7.5, generator=None, eta=0.0, deepcache=None, forced_latents=None, skip_timesteps=None,
            guidance_rescale=0.0, context_edit=None):
    """Returns dict(latents=final, per_step=[latents after each EXECUTED step], x0=[x0 preds], timesteps=[...],
    timesteps_run=[timesteps of the executed steps]).

    ``forced_latents[i]`` (if given) replaces the loop's latents before step i (teacher forcing).
    ``skip_timesteps``: loop INDICES that are skipped entirely -- no UNet call, no scheduler step, so a multistep
    scheduler's own step counter falls behind the grid exactly as in the reference
    (/root/reference/src/models.py:1220-1223,1338-1340, ``StableDiffusionModelSkipTimesteps``)."""
    skip = set(skip_timesteps) if skip_timesteps is not None else set()
    do_cfg = guidance_scale > 1
    ctx = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
    device = latents.device
    scheduler.set_timesteps(num_inference_steps, device=device)
    timesteps = scheduler.timesteps
    t_list = [int(t) for t in timesteps.tolist()]
    latents = latents * scheduler.init_noise_sigma
    extra = _extra(scheduler, generator, eta)
    per_step, x0s, ran = [], [], []
    if deepcache is not None:
        deepcache.reset()
    for i, t in enumerate(timesteps):
        if i in skip:
            continue
        ran.append(int(t))
        if forced_latents is not None:
            latents = forced_latents[i]
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        x_in = scheduler.scale_model_input(x_in, t)
        if deepcache is not None:
            noise_pred = deepcache.forward(x_in, t, ctx, t_list.index(int(t)))
        else:
            noise_pred = unet(x_in, t, encoder_hidden_states=ctx)[0]
        if do_cfg:
            noise_pred = _guide(noise_pred, guidance_scale, guidance_rescale)
        step = scheduler.step(noise_pred, t, latents, **extra, return_dict=False)
        latents = step[0]
        if len(step) == 2:
            x0s.append(step[1][0].unsqueeze(0))
        per_step.append(latents)
        ctx = _edited(context_edit, i, ctx)
    return dict(latents=latents, per_step=per_step, x0=x0s, timesteps=t_list, timesteps_run=ran)


def switch_timestamp(timesteps_first, timesteps_second, num_step_switch, type_switch="closest"):
    """/root/reference/src/models.py:704-730."""
    first = list(timesteps_first[:num_step_switch].cpu().numpy())
    if type_switch == "closest":
        dist = [abs(t - first[-1]) for t in timesteps_second.cpu()]
        second = list(timesteps_second[np.argmin(dist):].cpu().numpy())
    elif type_switch == "left_closest":
        idx = [i for i, t in enumerate(timesteps_second.cpu()) if t - first[-1] >= 0]
        second = list(timesteps_second[idx[-1]:].cpu().numpy())
    else:
        idx = [i for i, t in enumerate(timesteps_second.cpu()) if t - first[-1] <= 0]
        second = list(timesteps_second[idx[0]:].cpu().numpy())
    return first, second


@torch.no_grad()
def denoise_two(unet, scheduler_first, scheduler_second, prompt_embeds, negative_prompt_embeds, latents,
                num_inference_steps_first, num_step_switch, type_switch="closest", guidance_scale=7.5,
                generator=None, eta=0.0, forced_latents=None, guidance_rescale=0.0, context_edit=None):
    do_cfg = guidance_scale > 1
    ctx = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
    device = latents.device
    scheduler_first.set_timesteps(num_inference_steps_first, device=device)
    ts1 = scheduler_first.timesteps
    scheduler_second.set_timesteps(device=device, timesteps=ts1.cpu().numpy())
    ts2 = scheduler_second.timesteps
    first, second = switch_timestamp(ts1, ts2, num_step_switch, type_switch)
    latents = latents * scheduler_first.init_noise_sigma
    e1, e2 = _extra(scheduler_first, generator, eta), _extra(scheduler_second, generator, eta)
    per_step = []
    for i, t in enumerate(first + second):
        sched, extra = (scheduler_first, e1) if i < len(first) else (scheduler_second, e2)
        if forced_latents is not None:
            latents = forced_latents[i]
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        noise_pred = unet(x_in, torch.as_tensor(t, device=device), encoder_hidden_states=ctx)[0]
        if do_cfg:
            noise_pred = _guide(noise_pred, guidance_scale, guidance_rescale)
        # the history seeding of models.py:603-611 cannot run (SURVEY C-4) and is a no-op for order <= 2
        latents = sched.step(noise_pred, t, latents, **extra, return_dict=False)[0]
        per_step.append(latents)
        ctx = _edited(context_edit, i, ctx)
    return dict(latents=latents, per_step=per_step, timesteps=([int(t) for t in first], [int(t) for t in second]))


def interleave_partition(timesteps_main, solver_order, interliving_steps):
    """/root/reference/src/models.py:944-961: main-grid steps are grouped ``solver_order`` at a time; in every
    group listed in ``interliving_steps`` the first timestep is handed to the inter scheduler and the rest are
    dropped.  Returns (kept timesteps, inter timesteps) as python ints."""
    kept, inter = [], []
    for i, t in enumerate(int(v) for v in timesteps_main):
        if i // solver_order in interliving_steps:
            if i % solver_order != 0:
                continue
            inter.append(t)
        kept.append(t)
    return kept, inter


@torch.no_grad()
def denoise_interleaved(unet, scheduler_main, scheduler_inter, prompt_embeds, negative_prompt_embeds, latents,
                        num_inference_steps, interliving_steps, guidance_scale=7.5, generator=None, eta=0.0,
                        guidance_rescale=0.0):
    """The interleaved loop as written: the main (multistep DPM) scheduler walks its grid with its OWN step counter
    (models.py:1035 passes ``t`` but the step index only ever increments, src/schedulers.py:176), the inter
    scheduler (set up on N // order steps, models.py:888-894) replaces whole groups, and after every step the
    OTHER scheduler's history is fed ``convert_model_output(noise_pred, sample=new latents)``
    (models.py:1024-1031, 1045-1053).  ``convert_model_output`` returns one tensor for the ``++`` types and a pair
    otherwise (SURVEY C-1), so the literal history entry is only well-formed for ``++``; here it is the converted
    model output in both cases (the evident intent)."""
    do_cfg = guidance_scale > 1
    ctx = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
    device = latents.device
    order = scheduler_main.config.solver_order
    scheduler_main.set_timesteps(num_inference_steps, device=device)
    scheduler_inter.set_timesteps(num_inference_steps // order, device=device)
    kept, inter = interleave_partition(scheduler_main.timesteps.tolist(), order, interliving_steps)
    latents = latents * scheduler_main.init_noise_sigma          # prepare_latents uses self.scheduler (sigma 1.0)
    e_main, e_inter = _extra(scheduler_main, generator, eta), _extra(scheduler_inter, generator, eta)
    per_step = []

    def feed(s, noise_pred, new_latents):
        m = s.convert_model_output(noise_pred, sample=new_latents)[0]
        for j in range(s.config.solver_order - 1):
            s.model_outputs[j] = s.model_outputs[j + 1]
        s.model_outputs[-1] = m

    for t in kept:
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        noise_pred = unet(x_in, torch.as_tensor(t, device=device), encoder_hidden_states=ctx)[0]
        if do_cfg:
            noise_pred = _guide(noise_pred, guidance_scale, guidance_rescale)
        if t in inter:
            latents = scheduler_inter.step(noise_pred, t, latents, **e_inter, return_dict=False)[0]
            feed(scheduler_main, noise_pred, latents)
        else:
            latents = scheduler_main.step(noise_pred, t, latents, **e_main, return_dict=False)[0]
            if isinstance(scheduler_inter, type(scheduler_main)):
                feed(scheduler_inter, noise_pred, latents)
        per_step.append(latents)
    return dict(latents=latents, per_step=per_step, timesteps=(kept, inter))
