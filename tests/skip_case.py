"""Teacher-forced run of the skip-timesteps pipeline (``stable_diffusion_model_skip_timesteps``,
/root/reference/src/models.py:1138-1467: the single-scheduler loop with ``if i in skip_timesteps: continue`` at
:1338-1340, indices from the driver's ``skip_steps`` lists, skip_steps_exp.py:55-62) against the oracle loop
(oracle/pipeline.py ``denoise(..., skip_timesteps=...)``).

Shared by tests/test_skip_steps_gpu.py (real engine, unit-variance fixture) and its CPU dry run in
tests/test_pipeline_host_cpu.py (fake engine over the tiny oracle UNet), so the harness itself -- index mapping
between loop indices and executed steps, forcing, callback contract -- is checked without a GPU.

A skipped index runs neither the UNet nor the scheduler, so the multistep solver's own step counter falls behind
the grid: executed step j feeds the UNet the grid's timestep ``t[executed[j]]`` while the solver integrates its
j-th sigma interval.  Every executed step is entered with the same forced latents in the oracle and the product.
"""
from __future__ import annotations

import torch

DPMPP = dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")


def forward_process_path(alphas_cumprod, timesteps, z0, noise):
    """x_t = sqrt(abar_t) z0 + sqrt(1 - abar_t) n for every grid timestep (tests/parity_lib.py ``forced_path``)."""
    out = []
    for t in timesteps:
        a = float(alphas_cumprod[int(t)])
        out.append(a ** 0.5 * z0 + (1 - a) ** 0.5 * noise)
    return out


def run(model, net, net16, pe, ne, z0, noise, n, skip, overrides=None, guidance_scale=7.5):
    """Returns dict(engine=[max-abs per executed step], torch_bf16=[...] or None, seen=[(loop index, timestep)],
    executed=[loop indices], ref=oracle result, xmax=...)."""
    from oracle import schedulers as O
    from oracle.pipeline import denoise
    from sonicdiffusionbayeslab_b200 import schedulers as S

    kw = dict(DPMPP if overrides is None else overrides)
    cfg = O.SD15_SCHEDULER_CONFIG
    skip = [int(i) for i in skip]
    probe = O.DPMSolverScheduler.from_config(cfg, **kw)
    probe.set_timesteps(n)
    ts = [int(t) for t in probe.timesteps.tolist()]
    forced = forward_process_path(probe.alphas_cumprod, ts, z0, noise)
    executed = [i for i in range(n) if i not in set(skip)]
    first = forced[executed[0]]
    ref = denoise(net, O.DPMSolverScheduler.from_config(cfg, **kw), pe, ne, first, n, guidance_scale=guidance_scale,
                  forced_latents=forced, skip_timesteps=skip)
    floor = None
    if net16 is not None:                                      # stock PyTorch bf16, forced the same way
        floor = denoise(net16, O.DPMSolverScheduler.from_config(cfg, **kw), pe.bfloat16(), ne.bfloat16(),
                        first.bfloat16(), n, guidance_scale=guidance_scale,
                        forced_latents=[f.bfloat16() for f in forced], skip_timesteps=skip)["per_step"]
    assert len(ref["per_step"]) == len(executed) and ref["timesteps_run"] == [ts[i] for i in executed]

    model.scheduler = S.DPMSolverScheduler.from_config(cfg, **kw)
    errs, seen = [], []

    def cb(pipe, i, t, kwargs):
        j = len(seen)                                          # executed-step counter; ``i`` is the LOOP index
        seen.append((int(i), int(t)))
        errs.append((kwargs["latents"].float() - ref["per_step"][j].float()).abs().max().item())
        if j + 1 < len(executed):
            return {"latents": forced[executed[j + 1]].to(kwargs["latents"].dtype)}
        return {}

    out, secs, x0 = model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=first, num_inference_steps=n,
                          guidance_scale=guidance_scale, output_type="latent", skip_timesteps=skip,
                          callback_on_step_end=cb)
    res = dict(engine=errs, seen=seen, executed=executed, ref=ref, timesteps=ts, out=out, secs=secs, x0=x0,
               xmax=max(r.abs().max().item() for r in ref["per_step"]), torch_bf16=None)
    if floor is not None:
        res["torch_bf16"] = [(f.float() - r.float()).abs().max().item() for f, r in zip(floor, ref["per_step"])]
    return res
