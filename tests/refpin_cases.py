"""Cases of the reference-source pins, shared by ``tests/golden/make_reference_pins.py`` (which runs them
through the REFERENCE'S OWN source, oracle/refexec.py, and writes ``tests/golden/reference_pins.npz`` /
``reference_pins.json``) and by the tests (which run them through the oracle and the product).

Every input is derived from seeds, so the fixtures hold outputs only.  Scheduler cases feed a synthetic
epsilon sequence ``eps_i = 0.7 z_i + 0.2 x_i`` (no UNet); pipeline cases run a TINY oracle UNet
(3.2 M parameters, latent 8x8) so the whole set takes seconds on a CPU -- the pins are about the loop /
scheduler logic the reference owns, not about the network.
"""
from __future__ import annotations

import torch

SD15 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
            trained_betas=None, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1, clip_sample=False,
            prediction_type="epsilon", timestep_spacing="leading")

B, C, HW = 3, 4, 8          # B = 3: the unpatched ``++`` step must raise (it only unpacks for B == 2, SURVEY C-1)

# name -> (scheduler kind, config overrides, num_inference_steps, needs the C-1 source patch, noise seed or None)
SCHEDULER_CASES = {
    "dpm_o1_n10": ("dpm", dict(solver_order=1, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 10, False, None),
    "dpm_o2_n10": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 10, False, None),
    "dpm_o2_n20_heun": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min",
                                     solver_type="heun"), 20, False, None),
    "dpm_o3_n10": ("dpm", dict(solver_order=3, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 10, False, None),
    "dpm_o3_n20": ("dpm", dict(solver_order=3, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 20, False, None),
    "sde_dpm_o2_n10": ("dpm", dict(solver_order=2, algorithm_type="sde-dpmsolver", final_sigmas_type="sigma_min"), 10,
                       False, 11),
    "dpmpp_o2_n25": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero"), 25, True, None),
    "dpmpp_o3_n20": ("dpm", dict(solver_order=3, algorithm_type="dpmsolver++", final_sigmas_type="zero"), 20, True, None),
    "dpmpp_o2_n8_heun": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", solver_type="heun"), 8, True, None),
    "sde_dpmpp_o2_n10": ("dpm", dict(solver_order=2, algorithm_type="sde-dpmsolver++"), 10, True, 11),
    # the other ``prediction_type`` branches of convert_model_output (schedulers.py:43-51, :72-79): the synthetic
    # sequence then plays the network's x0 / v / flow output
    "dpm_o2_n10_sample": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min",
                                       prediction_type="sample"), 10, False, None),
    "dpm_o2_n10_vpred": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min",
                                      prediction_type="v_prediction"), 10, False, None),
    "sde_dpm_o2_n8_vpred": ("dpm", dict(solver_order=2, algorithm_type="sde-dpmsolver", final_sigmas_type="sigma_min",
                                         prediction_type="v_prediction"), 8, False, 11),
    "dpmpp_o2_n10_sample": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", prediction_type="sample"), 10,
                            True, None),
    "dpmpp_o2_n10_vpred": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", prediction_type="v_prediction"),
                           10, True, None),
    "dpmpp_o3_n12_vpred": ("dpm", dict(solver_order=3, algorithm_type="dpmsolver++", prediction_type="v_prediction"),
                           12, True, None),
    "dpmpp_o2_n10_flow": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", prediction_type="flow_prediction"),
                          10, True, None),
}

# convert_model_output's thresholding branches (schedulers.py:58-59, :85-90) over the oracle's restatement of
# diffusers' ``_threshold_sample`` (product: the quantile kernel sonic_x0_threshold + sonic_latent_update_post)
THRESHOLD_CASES = {
    "dpm_o2_n10_thr": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min",
                                    thresholding=True, sample_max_value=2.0), 10, False, None),
    "dpmpp_o2_n10_thr": ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", thresholding=True,
                                      sample_max_value=2.0), 10, True, None),
}

ALL_SCHEDULER_CASES = {**SCHEDULER_CASES, **THRESHOLD_CASES}


def step_tolerance(name, dtype):
    """Teacher-forced per-step max-abs gate of the fused scheduler step against the reference-source fixtures
    (BASELINE.json north_star): fp32 I/O 1e-4; bf16 I/O 2e-2 of max(1, |tensor|max).  Thresholded cases in bf16: 4e-2 --
    the thresholded x0 lives in [-1, 1] but is formed from an x0 of magnitude ~10-20 whose bf16 rounding (and the bf16
    rounding of eps times sigma/alpha ~ 14 at the first step) is divided by s ~ 2, i.e. the error is set by the range
    BEFORE thresholding; the reference's own formulas evaluated on bf16 tensors are at 4e-2 ... 2e-1 there."""
    if dtype == torch.float32:
        return 1e-4
    return 4e-2 if name in THRESHOLD_CASES else 2e-2


# name -> dict(pipe=..., ...) ; "patch": needs the C-1 source patch
PIPELINE_CASES = {
    "loop_ddim6": dict(pipe="single", sched=("ddim", {}), steps=6, guidance=7.5, patch=False),
    "loop_pndm6": dict(pipe="single", sched=("pndm", {}), steps=6, guidance=7.5, patch=False),
    "loop_lcm4_rng": dict(pipe="single", sched=("lcm", {}), steps=4, guidance=0.0, patch=False, gen_seed=29,
                          draw_latents=True),
    "loop_dpm8": dict(pipe="single", sched=("dpm", dict(solver_order=2, algorithm_type="dpmsolver",
                                                         final_sigmas_type="sigma_min")), steps=8, guidance=7.5,
                      patch=False),
    "loop_dpmpp10": dict(pipe="single", sched=("dpm", dict(solver_order=2, algorithm_type="dpmsolver++",
                                                            final_sigmas_type="zero")), steps=10, guidance=7.5,
                         patch=True),
    # The second scheduler must accept ``set_timesteps(timesteps=...)`` (models.py:490-494): of the reference's
    # schedulers only DPM-Solver does, and with the reference's OWN subclass there the phase-1 history seeding
    # (models.py:603-611) raises (SURVEY C-4, pinned under "raises").  "dpm_stock" = the stock
    # DPMSolverMultistepScheduler stand-in (not an instance of the reference subclass -> no seeding).
    "two_ddim_dpmstock": dict(pipe="two", first=("ddim", {}), second=("dpm_stock", dict(algorithm_type="dpmsolver++")),
                              n1=10, k=3, type_switch="closest", guidance=7.5, patch=False),
    "two_dpm_dpmstock_left": dict(pipe="two", first=("dpm", dict(algorithm_type="dpmsolver",
                                                                  final_sigmas_type="sigma_min")),
                                  second=("dpm_stock", dict(algorithm_type="dpmsolver++")), n1=10, k=4,
                                  type_switch="left_closest", guidance=7.5, patch=False),
    "two_dpmpp_dpmstock_right": dict(pipe="two", first=("dpm", dict(algorithm_type="dpmsolver++")),
                                     second=("dpm_stock", dict(algorithm_type="dpmsolver", final_sigmas_type="sigma_min")),
                                     n1=12, k=5, type_switch="right_closest", guidance=7.5, patch=True),
    "inter_dpmpp_ddim": dict(pipe="inter", main=("dpm", dict(solver_order=2, algorithm_type="dpmsolver++")),
                             inter=("ddim", {}), steps=10, groups=[1, 3], guidance=7.5, patch=True),
    "skip_ddim": dict(pipe="skip", sched=("ddim", {}), steps=8, skip=[2, 5], guidance=7.5, patch=False),
    "skip_dpmpp": dict(pipe="skip", sched=("dpm", dict(solver_order=2, algorithm_type="dpmsolver++",
                                                        final_sigmas_type="zero")), steps=10, skip=[1, 2, 7],
                       guidance=7.5, patch=True),
    # ``guidance_rescale`` > 0: the ``rescale_noise_cfg`` branch of every loop (models.py:244-250, 583-589, 1001-1007)
    "loop_ddim6_rescale": dict(pipe="single", sched=("ddim", {}), steps=6, guidance=7.5, rescale=0.7, patch=False),
    "two_ddim_dpmstock_rescale": dict(pipe="two", first=("ddim", {}),
                                      second=("dpm_stock", dict(algorithm_type="dpmsolver++")), n1=10, k=3,
                                      type_switch="closest", guidance=7.5, rescale=0.5, patch=False),
    # ``callback_on_step_end_tensor_inputs`` beyond latents (models.py:263-273): the callback is handed the UNet context
    # (``prompt_embeds`` = cat([negative, positive]) under guidance, models.py:154-155) and the negative embeddings,
    # and the ``prompt_embeds`` it returns after loop index ``at`` replace the context of every later step
    "loop_ddim6_ctxedit": dict(pipe="single", sched=("ddim", {}), steps=6, guidance=7.5, patch=False,
                               ctx_edit=dict(at=2, scale=0.5)),
    "two_ddim_dpmstock_ctxedit": dict(pipe="two", first=("ddim", {}),
                                      second=("dpm_stock", dict(algorithm_type="dpmsolver++")), n1=10, k=3,
                                      type_switch="closest", guidance=7.5, patch=False, ctx_edit=dict(at=4, scale=-1.0)),
    # the deprecated ``callback`` / ``callback_steps`` pair of the single-scheduler loops (models.py:275-282): PLMS has
    # N + 1 timesteps, so one warm-up index is passed over (models.py:205)
    "loop_pndm6_legacy_cb": dict(pipe="single", sched=("pndm", {}), steps=6, guidance=7.5, patch=False,
                                 legacy_cb=dict(callback_steps=2)),
    # ``num_images_per_prompt`` > 1: latents for B x n images, each prompt's embeddings repeated n times in place
    # (models.py:123-128,139-149,173; the repetition itself is diffusers' ``encode_prompt``, restated in the stub)
    "loop_ddim6_n2": dict(pipe="single", sched=("ddim", {}), steps=6, guidance=7.5, patch=False, n_img=2, gen_seed=31,
                          draw_latents=True),
    "inter_dpmpp_ddim_n2": dict(pipe="inter", main=("dpm", dict(solver_order=2, algorithm_type="dpmsolver++")),
                                inter=("ddim", {}), steps=10, groups=[1, 3], guidance=7.5, patch=True, n_img=2,
                                gen_seed=37, draw_latents=True),
    "inter_dpmpp_ddim_rescale": dict(pipe="inter", main=("dpm", dict(solver_order=2, algorithm_type="dpmsolver++")),
                                     inter=("ddim", {}), steps=10, groups=[1, 3], guidance=7.5, rescale=1.0,
                                     patch=True),
}

# (n_first, n_second or None = second runs on the first grid like the pipeline, num_step_switch)
SWITCH_CASES = [(10, None, 3), (10, None, 5), (20, None, 5), (20, None, 10), (30, None, 5), (30, None, 10),
                (10, 7, 3), (20, 13, 10), (25, 50, 7), (50, 9, 20), (10, 10, 1), (10, 10, 10)]
SWITCH_TYPES = ("closest", "left_closest", "right_closest")


def tiny_unet(seed=29):
    from oracle.unet import UNetConfig, make_unet

    cfg = UNetConfig(sample_size=HW, block_out_channels=(32, 64, 64, 64), num_heads=4, cross_attention_dim=32,
                     norm_num_groups=8)
    return make_unet(seed, cfg)


def pipeline_inputs(seed=1):
    g = torch.Generator().manual_seed(seed)
    pe, ne = torch.randn(B, 7, 32, generator=g), torch.randn(B, 7, 32, generator=g)
    lat = torch.randn(B, C, HW, HW, generator=g)
    return pe, ne, lat


def make_scheduler(kind, overrides, *, module=None, ref=None):
    """``module``: oracle.schedulers or the product's schedulers; ``ref``: a refexec namespace (reference source)."""
    if ref is not None:
        from oracle import schedulers as O

        cls = {"dpm": ref.DPMSolverScheduler, "ddim": ref.DDIMSchedulerMy, "lcm": ref.LCMScheduler,
               "pndm": O.PNDMScheduler,             # PNDM is stock diffusers in the reference (no source of its own)
               "dpm_stock": O.DPMSolverScheduler}[kind]
    else:
        cls = {"dpm": module.DPMSolverScheduler, "dpm_stock": module.DPMSolverScheduler,
               "ddim": getattr(module, "DDIMSchedulerMy", None) or module.DDIMScheduler, "lcm": module.LCMScheduler,
               "pndm": module.PNDMScheduler}[kind]
    return cls.from_config(SD15, **overrides)


def run_scheduler_case(sched, n_steps, noise_seed, device="cpu", dtype=torch.float32, teacher=None):
    """Drives ``sched.step`` over the synthetic epsilon sequence.  Returns (prev_samples, x0_preds, timesteps).
    ``teacher``: list of prev_samples to continue from (teacher forcing, so kernel errors do not compound)."""
    sched.set_timesteps(n_steps, device=device)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, HW, HW, generator=g).to(device=device, dtype=dtype)
    gen = None
    if noise_seed is not None:
        gen = torch.Generator().manual_seed(noise_seed)      # CPU generator: randn_tensor draws on the host
    prevs, x0s = [], []
    for i, t in enumerate(sched.timesteps):
        eps = (0.7 * torch.randn(B, C, HW, HW, generator=g).to(device) + 0.2 * x.float()).to(dtype)
        kw = {"generator": gen} if gen is not None else {}
        out = sched.step(eps, t, x, return_dict=False, **kw)
        prevs.append(out[0])
        x0s.append(out[1])
        x = out[0] if teacher is None else teacher[i].to(device=device, dtype=dtype)
    return prevs, x0s, [int(t) for t in sched.timesteps.tolist()]


CB_INPUTS = ["latents", "prompt_embeds", "negative_prompt_embeds"]


def context_edit(case):
    """The deterministic edit of the ``ctx_edit`` cases: after loop index ``at`` the context rows are reversed (under
    guidance that swaps the negative and the positive halves) and scaled."""
    e = case.get("ctx_edit")
    if not e:
        return None

    def edit(i, ctx):
        return ctx.flip(0) * e["scale"] if i == e["at"] else None

    return edit


class EditingRecorder:
    """``callback_on_step_end`` for the ``ctx_edit`` cases: records like ``refexec.Recorder`` plus the shapes of the
    tensors it is handed, and returns the edited ``prompt_embeds``."""

    def __init__(self, edit):
        self.per_step, self.timesteps, self.shapes, self.edit = [], [], [], edit

    def __call__(self, pipe, i, t, kwargs):
        self.per_step.append(kwargs["latents"].clone())
        self.timesteps.append(int(t))
        self.shapes.append({k: tuple(v.shape) for k, v in kwargs.items()})
        new = self.edit(i, kwargs["prompt_embeds"])
        return {} if new is None else {"prompt_embeds": new}


class LegacyCallback:
    """The deprecated ``callback(step_idx, t, latents)``: records (step_idx, t) and a checksum of the latents."""

    def __init__(self):
        self.calls = []

    def __call__(self, step_idx, t, latents):
        self.calls.append([int(step_idx), int(t), round(float(latents.double().abs().sum()), 6)])


def run_pipeline_reference(case, ns, net):
    """One pipeline case through the reference's own ``call`` (models.py) compiled by oracle/refexec.py.
    Returns dict(per_step=[latents after each executed step], timesteps=[...], final=latents, n_x0=int,
    num_timesteps=int)."""
    from oracle import refexec
    from oracle import schedulers as O

    pe, ne, lat = pipeline_inputs()
    rec = EditingRecorder(context_edit(case)) if case.get("ctx_edit") else refexec.Recorder()
    default = O.PNDMScheduler.from_config(SD15)            # what from_pretrained leaves in pipe.scheduler
    kind = case["pipe"]
    common = dict(prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=case["guidance"], output_type="pt",
                  callback_on_step_end=rec)
    if case.get("ctx_edit"):
        common["callback_on_step_end_tensor_inputs"] = list(CB_INPUTS)
    legacy = None
    if case.get("legacy_cb"):
        legacy = LegacyCallback()
        common.update(callback=legacy, callback_steps=case["legacy_cb"]["callback_steps"])
    if case.get("gen_seed") is not None:
        common["generator"] = torch.Generator().manual_seed(case["gen_seed"])
    if case.get("rescale"):
        common["guidance_rescale"] = case["rescale"]
    if case.get("n_img"):
        common["num_images_per_prompt"] = case["n_img"]
    if not case.get("draw_latents"):
        common["latents"] = lat
    else:
        common["height"], common["width"] = 8 * HW, 8 * HW
    if kind in ("single", "skip"):
        cls = ns.StableDiffusionModel if kind == "single" else ns.StableDiffusionModelSkipTimesteps
        pipe = cls(net, scheduler=make_scheduler(*case["sched"], ref=ns))
        kw = dict(num_inference_steps=case["steps"])
        if kind == "skip":
            kw["skip_timesteps"] = list(case["skip"])
        out, _, x0 = pipe(**common, **kw)
    elif kind == "two":
        # models.py:638 reads ``self.scheduler.config.solver_order`` after every step: with the stock PNDM default
        # that from_pretrained leaves there (two_schedulers.py:44-62 never replaces it) the call raises at i = 0
        # (pinned under "raises"); the harness parks a DPM-Solver config there so the loop itself can run.
        if not case.get("pndm_default"):
            default = O.DPMSolverScheduler.from_config(SD15)
        pipe = ns.StableDiffusionModelTwoSchedulers(net, scheduler=default)
        pipe.scheduler_first = make_scheduler(*case["first"], ref=ns)
        pipe.scheduler_second = make_scheduler(*case["second"], ref=ns)
        out, _, x0 = pipe(**common, num_inference_steps_first=case["n1"], num_inference_steps_second=case["n1"],
                          num_step_switch=case["k"], type_switch=case["type_switch"])
    else:
        pipe = ns.StableDiffusionModelInterlivingSchedulers(net, scheduler=default)
        pipe.scheduler_main = make_scheduler(*case["main"], ref=ns)
        pipe.scheduler_inter = make_scheduler(*case["inter"], ref=ns)
        out, _, x0 = pipe(**common, num_inference_steps=case["steps"], interliving_steps=list(case["groups"]))
    return dict(per_step=rec.per_step, timesteps=rec.timesteps, final=out.images, n_x0=len(x0),
                num_timesteps=pipe.num_timesteps, cb_shapes=getattr(rec, "shapes", None),
                legacy_calls=legacy.calls if legacy else None)


def run_pipeline_oracle(case, net):
    """The same case through the oracle restatement (oracle/pipeline.py)."""
    from oracle import pipeline as P
    from oracle import schedulers as O

    pe, ne, lat = pipeline_inputs()
    gen = torch.Generator().manual_seed(case["gen_seed"]) if case.get("gen_seed") is not None else None
    kind = case["pipe"]
    n_img = case.get("n_img", 1)
    if n_img > 1:                                            # the n images of a prompt are adjacent rows
        pe, ne = pe.repeat_interleave(n_img, dim=0), ne.repeat_interleave(n_img, dim=0)
    if case.get("draw_latents"):
        lat = P.prepare_latents((B * n_img, C, HW, HW), gen, "cpu", pe.dtype)
    rs = case.get("rescale", 0.0)
    if kind == "single":
        r = P.denoise(net, make_scheduler(*case["sched"], module=O), pe, ne, lat, case["steps"],
                      guidance_scale=case["guidance"], generator=gen, guidance_rescale=rs,
                      context_edit=context_edit(case))
        ts = r["timesteps"]
    elif kind == "skip":
        r = P.denoise(net, make_scheduler(*case["sched"], module=O), pe, ne, lat, case["steps"],
                      guidance_scale=case["guidance"], generator=gen, skip_timesteps=case["skip"], guidance_rescale=rs)
        ts = r["timesteps_run"]
    elif kind == "two":
        r = P.denoise_two(net, make_scheduler(*case["first"], module=O), make_scheduler(*case["second"], module=O),
                          pe, ne, lat, case["n1"], case["k"], case["type_switch"], guidance_scale=case["guidance"],
                          guidance_rescale=rs, context_edit=context_edit(case))
        ts = r["timesteps"][0] + r["timesteps"][1]
    else:
        r = P.denoise_interleaved(net, make_scheduler(*case["main"], module=O), make_scheduler(*case["inter"], module=O),
                                  pe, ne, lat, case["steps"], case["groups"], guidance_scale=case["guidance"],
                                  guidance_rescale=rs)
        ts = r["timesteps"][0]
    return dict(per_step=r["per_step"], timesteps=[int(t) for t in ts], final=r["latents"])
