"""Shared machinery of the ABSOLUTE-tolerance parity tests (tests/test_parity_abs_gpu.py) and of
``tools/parity_report.py`` (which writes the measured table to profiles/).

BASELINE.json's north_star states the latent tolerance as an absolute number (max-abs <= 2e-2 in bf16 per step).
That is only meaningful when the trajectory has Stable-Diffusion-like magnitudes (|x| of a few units, eps of unit
variance); the plain random-init oracle UNet drives |x| to 40-80, where one bf16 ulp is already 0.25.  The
``unit-variance fixture`` below therefore rescales ``conv_out`` of the seeded oracle UNet so that eps has unit
variance on N(0,1) latents -- ONE state dict, shared by the fp32 oracle, stock-PyTorch bf16 and the engine.
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
    g = torch.Generator(device=device).manual_seed(seed)
    pe = torch.randn(B, 77, 768, device=device, generator=g).bfloat16().float()
    ne = torch.randn(B, 77, 768, device=device, generator=g).bfloat16().float()
    lat = torch.randn(B, 4, 64, 64, device=device, generator=g)
    return pe, ne, lat


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


def teacher_forced(name, net, net16, sd, device, io_dtype=torch.bfloat16, B=2, steps=None, model=None):
    """Per-step ABSOLUTE max-abs error of the latents leaving each step, every implementation teacher-forced from
    the fp32 oracle's latents entering that step.  Returns dict(engine=[...], torch_bf16=[...], xmax=[...],
    timesteps=[...])."""
    from oracle import schedulers as O
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise, denoise_two
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    kind, pcls, ocls, kw, n, extra = CASES[name]
    n = steps or n
    cfg = O.SD15_SCHEDULER_CONFIG
    pe, ne, lat = inputs(device, B)
    errs = []

    def cb_factory(ref_steps):
        def cb(pipe, i, t, kwargs):
            want = ref_steps[i]
            errs.append((kwargs["latents"].float() - want.float()).abs().max().item())
            return {"latents": want.to(kwargs["latents"].dtype)}
        return cb

    if kind == "single":
        dc = dc16 = None
        if "deepcache" in extra:
            dc, dc16 = DeepCacheOracle(net), DeepCacheOracle(net16)
            for d in (dc, dc16):
                d.set_params(cache_interval=extra["deepcache"][0], cache_branch_id=extra["deepcache"][1])
        ref = denoise(net, getattr(O, ocls).from_config(cfg, **kw), pe, ne, lat, n, deepcache=dc)
        forced = [lat] + ref["per_step"][:-1]
        floor = denoise(net16, getattr(O, ocls).from_config(cfg, **kw), pe.bfloat16(), ne.bfloat16(), lat.bfloat16(), n,
                        deepcache=dc16, forced_latents=[f.bfloat16() for f in forced])["per_step"]
        model = model or make_model(sd, device, io_dtype=io_dtype)
        model.scheduler = getattr(S, pcls).from_config(cfg, **kw)
        helper = None
        if "deepcache" in extra:
            helper = DeepCacheSDHelper(pipe=model)
            helper.set_params(cache_interval=extra["deepcache"][0], cache_branch_id=extra["deepcache"][1])
            helper.enable()
        model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, num_inference_steps=n, guidance_scale=7.5,
              output_type="latent", callback_on_step_end=cb_factory(ref["per_step"]))
        if helper:
            helper.disable()
        assert model.scheduler.timesteps.tolist() == ref["timesteps"]
        ts = ref["timesteps"]
    else:
        k = extra["k"]
        ref = denoise_two(net, O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg), pe, ne, lat, n, k)
        floor = None                                  # denoise_two has no teacher forcing: engine vs fp32 oracle only
        model = model or make_model(sd, device, M.StableDiffusionModelTwoSchedulers, io_dtype=io_dtype)
        model.scheduler_first = S.DDIMSchedulerMy.from_config(cfg)
        model.scheduler_second = S.DPMSolverScheduler.from_config(cfg)
        model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, num_inference_steps_first=n,
              num_inference_steps_second=n, num_step_switch=k, guidance_scale=7.5, output_type="latent",
              callback_on_step_end=cb_factory(ref["per_step"]))
        first, second = model.last_timesteps
        assert ([int(t) for t in first], [int(t) for t in second]) == ref["timesteps"]
        ts = ref["timesteps"][0] + ref["timesteps"][1]
    out = dict(engine=errs, xmax=[r.abs().max().item() for r in ref["per_step"]], timesteps=ts)
    if floor is not None:
        out["torch_bf16"] = [(f.float() - r).abs().max().item() for f, r in zip(floor, ref["per_step"])]
    return out
