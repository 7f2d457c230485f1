"""Loop-only images/s of the five BASELINE.json configurations on ONE B200 (per-GPU share of each config), through
the plugin call.  Not the bench.py contract line (that is configs[1]); printed for DESIGN.md / profiles."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sonicdiffusionbayeslab_b200 import models as M
from sonicdiffusionbayeslab_b200 import schedulers as S
from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

dev = torch.device("cuda:0")
CFG = M.SD15_SCHEDULER_CONFIG


def run(name, model, batch, reps=2, **kw):
    g = torch.Generator().manual_seed(29)
    pe = torch.randn(batch, 77, 768, generator=g).bfloat16().to(dev)
    ne = torch.randn(batch, 77, 768, generator=g).bfloat16().to(dev)
    lat = torch.randn(batch, 4, 64, 64, generator=g).bfloat16().to(dev)
    secs = []
    for _ in range(reps + 1):
        _, s, _ = model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, output_type="latent",
                        generator=torch.Generator().manual_seed(1), **kw)
        secs.append(s)
    s = min(secs[1:])
    kinds = getattr(model, "last_step_kinds", [])
    out = {"config": name, "batch_per_gpu": batch, "unet_evals": model.num_timesteps, "loop_s": round(s, 4),
           "images_per_s_per_gpu": round(batch / s, 2), "full_steps": kinds.count("full") or None,
           "cached_steps": kinds.count("cached") or None}
    print(json.dumps(out), flush=True)


base = M.StableDiffusionModel.from_pretrained("runwayml/stable-diffusion-v1-5", torch_dtype=torch.bfloat16).to(dev)
base.scheduler = S.DDIMSchedulerMy.from_config(CFG)
run("C1 ddim_config: DDIM 20 steps, batch 1, CFG 7.5", base, 1, num_inference_steps=20, guidance_scale=7.5)
base.scheduler = S.DPMSolverScheduler.from_config(CFG, solver_order=2, algorithm_type="dpmsolver++",
                                                  final_sigmas_type="zero")
run("C2 dpm_solver_config: DPM-Solver++ 25 steps, batch 16, CFG 7.5", base, 16, num_inference_steps=25,
    guidance_scale=7.5)
base.scheduler = S.DDIMSchedulerMy.from_config(CFG)
h = DeepCacheSDHelper(pipe=base)
h.set_params(cache_interval=3, cache_branch_id=0)
h.enable()
run("C3 deep_cache_config: DDIM 50 steps + DeepCache interval 3, batch 8 (64 over 8 GPUs), CFG 7.5", base, 8,
    num_inference_steps=50, guidance_scale=7.5)
h.disable()
base.scheduler = S.LCMScheduler.from_config(CFG)
run("C4 consistency_model_config: LCM 4 steps, batch 16 (128 over 8 GPUs), guidance 0", base, 16,
    num_inference_steps=4, guidance_scale=0.0)
run("C4 consistency_model_config: LCM 4 steps, batch 64 (128 over 2 GPUs), guidance 0", base, 64,
    num_inference_steps=4, guidance_scale=0.0)
two = M.StableDiffusionModelTwoSchedulers(base._unet_sd, base.vae, base.text_encoder, base.tokenizer, base.scheduler)
two.to(dev)
two._weights = base._weights
two.scheduler_first = S.DDIMSchedulerMy.from_config(CFG)
two.scheduler_second = S.DPMSolverScheduler.from_config(CFG, solver_order=2, algorithm_type="dpmsolver++",
                                                        final_sigmas_type="zero")
run("C5 two_schedulers_config: DDIM -> DPM-Solver++ switch, N1=20 k=10, batch 32, CFG 7.5", two, 32,
    num_inference_steps_first=20, num_inference_steps_second=20, num_step_switch=10, guidance_scale=7.5)
