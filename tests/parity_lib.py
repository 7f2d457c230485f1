"""Shared machinery of the ABSOLUTE-tolerance parity tests (tests/test_parity_abs_gpu.py) and of
``tools/parity_report.py`` (which writes the measured table to profiles/).

BASELINE.json's north_star states the latent tolerance as an absolute number (max-abs <= 2e-2 in bf16 per step).
That is only meaningful when the trajectory has Stable-Diffusion-like magnitudes (|x| of a few units, eps of unit
variance); a random-init UNet's own trajectory reaches |x| of 40-80, where one bf16 ulp is already 0.25.  The
fixture therefore has three parts:
  * ``unit_variance_unet``: ``conv_out`` of the seeded oracle UNet rescaled so eps has unit variance on N(0,1)
    latents -- ONE state dict, shared by the fp32 oracle, stock-PyTorch bf16 and the engine;
  * ``inputs``: a negative prompt embedding close to the positive one (correlated cond / uncond predictions, as for
    a trained model; independent random contexts make classifier-free guidance 7.5 inflate eps tenfold);
  * ``forced_path``: every step is entered with a sample of the FORWARD process
    x_t = sqrt(abar_t) z0 + sqrt(1 - abar_t) n (a random network's eps does not track the noise in x, so its own
    trajectory grows by sqrt(abar_0 / abar_T) ~ 15x).
"""
from __future__ import annotations

import copy

import torch


def unit_variance_unet(device, seed=29):
    """(fp32 oracle net, its bf16 copy, the calibration scale).  eps std ~ 1 at t = 981 / 501 / 21."""
    from oracle.unet import make_unet

    net = make_unet(seed).to(device)
    g = torch.Generator(device=device).manual_seed(1234)
    x = torch.randn(2, 4, 64, 64, device=device, generator=g)
    ctx = torch.randn(2, 77, 768, device=device, generator=g)
    with torch.no_grad():
        stds = [net(x, torch.tensor(t, device=device), encoder_hidden_states=ctx)[0].std().item() for t in (981, 501, 21)]
        scale = 1.0 / (sum(stds) / len(stds))
        net.conv_out.weight.mul_(scale)
        net.conv_out.bias.mul_(scale)
    net16 = copy.deepcopy(net).to(torch.bfloat16)
    return net, net16, scale


def inputs(device, B=2, seed=29):
    """Prompt / negative-prompt embeddings and the two N(0,1) tensors the forced trajectory is built from.  The
    negative embedding is a small perturbation of the positive one, so the conditional and unconditional
    predictions are strongly correlated as they are for a trained model -- with independent random contexts
    classifier-free guidance 7.5 alone inflates eps tenfold."""
    g = torch.Generator(device=device).manual_seed(seed)
    pe = torch.randn(B, 77, 768, device=device, generator=g).bfloat16().float()
    ne = (pe + 0.05 * torch.randn(B, 77, 768, device=device, generator=g)).bfloat16().float()
    z0 = torch.randn(B, 4, 64, 64, device=device, generator=g)
    n = torch.randn(B, 4, 64, 64, device=device, generator=g)
    return pe, ne, z0, n


def forced_path(alphas_cumprod, timesteps, z0, n):
    """Latents ENTERING each step: a sample path of the forward process, x_t = sqrt(a_t) z0 + sqrt(1 - a_t) n --
    what a trained model's trajectory looks like (|x| of a few units).  A random-init network's own trajectory
    grows by sqrt(a_0 / a_T) ~ 15x because its eps does not track the noise in x."""
    out = []
    for t in timesteps:
        a = float(alphas_cumprod[int(t)])
        out.append(a ** 0.5 * z0 + (1 - a) ** 0.5 * n)
    return out


def make_model(sd, device, cls=None, io_dtype=torch.bfloat16):
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer

    cls = cls or M.StableDiffusionModel
    m = cls(sd, vae=None, text_encoder=None, tokenizer=HashTokenizer(),
            scheduler=S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG), torch_dtype=io_dtype)
    m.device = torch.device(device)
    return m


CASES = {
    # name: (kind, product scheduler cls name, oracle cls name, overrides, steps, extra)
    "ddim20": ("single", "DDIMSchedulerMy", "DDIMScheduler", {}, 20, {}),
    "dpmpp25": ("single", "DPMSolverScheduler", "DPMSolverScheduler",
                dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero"), 25, {}),
    "pndm20": ("single", "PNDMScheduler", "PNDMScheduler", {}, 20, {}),
    "deepcache_ddim12_i3": ("single", "DDIMSchedulerMy", "DDIMScheduler", {}, 12, dict(deepcache=(3, 0))),
    "two_20_k10": ("two", None, None, {}, 20, dict(k=10)),
}


class StopAfter(Exception):
    pass


def teacher_forced(name, net, net16, sd, device, io_dtype=torch.bfloat16, B=2, max_steps=None, model=None):
    """Per-step ABSOLUTE max-abs error of the latents LEAVING each step.  Every implementation (fp32 oracle,
    stock-PyTorch bf16, engine) enters step i with the same forced latents (``forced_path``).  ``max_steps``:
    stop after that many steps of the genuine schedule.  Returns dict(engine=[...], torch_bf16=[...] or None,
    xmax=[|latents|max leaving each step], timesteps=[...])."""
    from oracle import schedulers as O
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise, denoise_two
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    kind, pcls, ocls, kw, n, extra = CASES[name]
    cfg = O.SD15_SCHEDULER_CONFIG
    pe, ne, z0, noise = inputs(device, B)
    errs = []

    # the first max_steps steps of the GENUINE schedule: later loop indices are skipped, the scheduler still knows the
    # full grid (order / lower-order-final decisions depend on its length)
    skip = None if max_steps is None else list(range(max_steps, 1000))

    def callback(ref_steps, forced):
        def cb(pipe, i, t, kwargs):
            errs.append((kwargs["latents"].float() - ref_steps[i].float()).abs().max().item())
            if i + 1 == len(ref_steps):
                if max_steps is not None:
                    raise StopAfter
                return {}
            return {"latents": forced[i + 1].to(kwargs["latents"].dtype)}
        return cb

    def run_product(fn):
        try:
            fn()
        except StopAfter:
            pass

    floor = None
    if kind == "single":
        probe = getattr(O, ocls).from_config(cfg, **kw)
        probe.set_timesteps(n)
        ts = [int(t) for t in probe.timesteps.tolist()]
        forced = forced_path(probe.alphas_cumprod, ts, z0, noise)
        dc = dc16 = None
        if "deepcache" in extra:
            dc = DeepCacheOracle(net)
            dc.set_params(cache_interval=extra["deepcache"][0], cache_branch_id=extra["deepcache"][1])
            if net16 is not None:
                dc16 = DeepCacheOracle(net16)
                dc16.set_params(cache_interval=extra["deepcache"][0], cache_branch_id=extra["deepcache"][1])
        ref = denoise(net, getattr(O, ocls).from_config(cfg, **kw), pe, ne, forced[0], n, deepcache=dc,
                      forced_latents=forced, skip_timesteps=skip)
        if net16 is not None:
            floor = denoise(net16, getattr(O, ocls).from_config(cfg, **kw), pe.bfloat16(), ne.bfloat16(),
                            forced[0].bfloat16(), n, deepcache=dc16, forced_latents=[f.bfloat16() for f in forced],
                            skip_timesteps=skip)["per_step"]
        model = model or make_model(sd, device, io_dtype=io_dtype)
        model.scheduler = getattr(S, pcls).from_config(cfg, **kw)
        helper = None
        if "deepcache" in extra:
            helper = DeepCacheSDHelper(pipe=model)
            helper.set_params(cache_interval=extra["deepcache"][0], cache_branch_id=extra["deepcache"][1])
            helper.enable()
        run_product(lambda: model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=forced[0], num_inference_steps=n,
                                  guidance_scale=7.5, output_type="latent",
                                  callback_on_step_end=callback(ref["per_step"], forced)))
        if helper:
            helper.disable()
        assert model.scheduler.timesteps.tolist()[:len(ref["per_step"])] == ts[:len(ref["per_step"])]
        scheds = [model.scheduler]
    else:
        k = extra["k"]
        s1, s2 = O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg)
        probe = denoise_two(lambda x, t, encoder_hidden_states=None: (torch.zeros_like(x),), s1, s2, pe, ne, z0, n, k)
        ts = probe["timesteps"][0] + probe["timesteps"][1]
        forced = forced_path(s1.alphas_cumprod, ts, z0, noise)
        ref = denoise_two(net, O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg), pe, ne, forced[0],
                          n, k, forced_latents=forced)
        model = model or make_model(sd, device, M.StableDiffusionModelTwoSchedulers, io_dtype=io_dtype)
        model.scheduler_first = S.DDIMSchedulerMy.from_config(cfg)
        model.scheduler_second = S.DPMSolverScheduler.from_config(cfg)
        run_product(lambda: model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=forced[0],
                                  num_inference_steps_first=n, num_inference_steps_second=n, num_step_switch=k,
                                  guidance_scale=7.5, output_type="latent",
                                  callback_on_step_end=callback(ref["per_step"], forced)))
        first, second = model.last_timesteps
        assert ([int(t) for t in first], [int(t) for t in second]) == ref["timesteps"]
        scheds = [model.scheduler_first, model.scheduler_second]
    for s_ in scheds:                                      # a StopAfter skips the pipeline's own reset
        s_.x0_rows, s_.skip_x0, s_.rng_rows = None, False, None
    out = dict(engine=errs, xmax=[r.abs().max().item() for r in ref["per_step"]], timesteps=ts, torch_bf16=None)
    if floor is not None:
        out["torch_bf16"] = [(f.float() - r).abs().max().item() for f, r in zip(floor, ref["per_step"])]
    return out
