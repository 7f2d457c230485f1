"""GPU parity of the sampling loop: fused scheduler steps and whole pipelines vs the oracle.

Tolerances (stated, per BASELINE.json north_star):
  * timestep / index schedules, DeepCache full/cached pattern, switch point: bit-exact;
  * scheduler update given the SAME epsilon (isolates the fused latent-update kernel):
        fp32 I/O  max-abs <= 1e-4      bf16 I/O  max-abs <= 2e-2 (x' is bf16-rounded, |x| <~ 5);
  * whole step through the bf16 engine UNet, teacher-forced from the oracle's latents, vs the fp32
    oracle UNet: max-abs <= TF_TOL.  The UNet itself is bf16-vs-fp32 (max-abs ~1.5e-2 on eps) and
    classifier-free guidance 7.5 amplifies the eps error ~7.5*sqrt(2)x, so the same comparison is
    also made for stock PyTorch bf16 (the reference's own library path) and printed beside it.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TF_TOL = 2e-2          # teacher-forced, per step, bf16 engine vs fp32 oracle, CFG 7.5: max-abs relative
                       # to max(1, |latents|max) -- a random-init UNet drives |latents| to 10-80, where
                       # an absolute 2e-2 is below one bf16 ulp (0.25 at 32-64)
FREE_TOL_REL = 0.04    # free-running final latents (no teacher forcing, errors compound over the trajectory): max-abs
                       # relative to the latent range; measured 1.5 % (DPM++ 10 steps, LCM 4 steps)


def _rel(got, ref):
    ref = ref.float()
    return (got.float() - ref).abs().max().item() / max(1.0, ref.abs().max().item())


@pytest.fixture(scope="module")
def world(cuda):
    from oracle.unet import make_unet
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer

    net = make_unet(29).to(cuda)
    sd = dict(net.state_dict())

    def make(cls=M.StableDiffusionModel, dtype=torch.bfloat16):
        sched = S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG)
        m = cls(sd, vae=None, text_encoder=None, tokenizer=HashTokenizer(), scheduler=sched, torch_dtype=dtype)
        m.device = cuda
        return m

    g = torch.Generator(device="cuda").manual_seed(29)
    B = 2
    pe = torch.randn(B, 77, 768, device=cuda, generator=g).bfloat16().float()
    ne = torch.randn(B, 77, 768, device=cuda, generator=g).bfloat16().float()
    lat = torch.randn(B, 4, 64, 64, device=cuda, generator=g)
    import copy

    net16 = copy.deepcopy(net).to(torch.bfloat16)      # stock PyTorch bf16 path (noise floor / LCM reference)
    return dict(net=net, net16=net16, make=make, pe=pe, ne=ne, lat=lat, B=B, dev=cuda)


def _sched_pairs():
    from oracle import schedulers as O
    from sonicdiffusionbayeslab_b200 import schedulers as S

    return {
        "ddim": (S.DDIMSchedulerMy, O.DDIMScheduler, {}, 20),
        "dpm++2": (S.DPMSolverScheduler, O.DPMSolverScheduler,
                   dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero"), 25),
        "dpm2": (S.DPMSolverScheduler, O.DPMSolverScheduler,
                 dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 10),
        "dpm++3": (S.DPMSolverScheduler, O.DPMSolverScheduler,
                   dict(solver_order=3, algorithm_type="dpmsolver++", final_sigmas_type="zero"), 20),
        "dpm++2heun": (S.DPMSolverScheduler, O.DPMSolverScheduler,
                       dict(solver_order=2, algorithm_type="dpmsolver++", solver_type="heun"), 8),
        "sde-dpm++2": (S.DPMSolverScheduler, O.DPMSolverScheduler,
                       dict(solver_order=2, algorithm_type="sde-dpmsolver++"), 10),
        "lcm": (S.LCMScheduler, O.LCMScheduler, {}, 4),
        "pndm": (S.PNDMScheduler, O.PNDMScheduler, {}, 12),
    }


@pytest.mark.parametrize("name", list(_sched_pairs()))
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_fused_scheduler_step_matches_oracle(cuda, name, dtype, tol):
    """Same epsilon sequence into both: isolates schedule + fused update kernel (no UNet)."""
    from oracle.schedulers import SD15_SCHEDULER_CONFIG

    P, Oc, kw, n = _sched_pairs()[name]
    ps, os_ = P.from_config(SD15_SCHEDULER_CONFIG, **kw), Oc.from_config(SD15_SCHEDULER_CONFIG, **kw)
    ps.set_timesteps(n, device=cuda)
    os_.set_timesteps(n, device=cuda)
    assert ps.timesteps.tolist() == os_.timesteps.tolist()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 4, 64, 64, device=cuda, generator=g).to(dtype)
    xp, xo = x.clone(), x.clone()
    gp, go = (torch.Generator(device="cuda").manual_seed(11) for _ in range(2))
    worst = 0.0
    for t in os_.timesteps:
        eps = (0.7 * torch.randn(2, 4, 64, 64, device=cuda, generator=g) + 0.2 * xo.float()).to(dtype)
        kwp = {"generator": gp} if name in ("lcm", "sde-dpm++2") else {}
        kwo = {"generator": go} if name in ("lcm", "sde-dpm++2") else {}
        rp = ps.step(eps, t, xp, **kwp)
        ro = os_.step(eps, t, xo, **kwo)
        assert len(rp) == len(ro)
        for a, b in zip(rp, ro):
            # scaled by the tensor's range: early-step x0 predictions reach |x0| ~ 50, where one
            # bf16 ulp is 0.25 (the oracle itself rounds every op to bf16 there)
            scale = max(1.0, b.float().abs().max().item()) if dtype == torch.bfloat16 else 1.0
            worst = max(worst, (a.float() - b.float()).abs().max().item() / scale)
        # teacher-force the product from the oracle's state so errors do not compound
        xp, xo = ro[0].clone(), ro[0]
    assert worst <= tol, f"{name} {dtype}: max-abs {worst:.3e} > {tol}"


def _torch_bf16_floor(world, Oc, kw, n, forced):
    """Stock PyTorch bf16 UNet (the reference's own kind of path), teacher-forced the same way."""
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG

    r = denoise(world["net16"], Oc.from_config(SD15_SCHEDULER_CONFIG, **kw), world["pe"].bfloat16(),
                world["ne"].bfloat16(), world["lat"].bfloat16(), n, forced_latents=[f.bfloat16() for f in forced])
    return r["per_step"]


@pytest.mark.parametrize("name", ["ddim", "dpm++2", "pndm"])
def test_pipeline_teacher_forced(world, name):
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG

    P, Oc, kw, n = _sched_pairs()[name]
    n = min(n, 8)
    ref = denoise(world["net"], Oc.from_config(SD15_SCHEDULER_CONFIG, **kw), world["pe"], world["ne"], world["lat"], n)
    forced = [world["lat"]] + ref["per_step"][:-1]          # oracle's latents entering step i
    model = world["make"]()
    model.scheduler = P.from_config(SD15_SCHEDULER_CONFIG, **kw)
    errs = []

    def cb(pipe, i, t, kwargs):
        errs.append(_rel(kwargs["latents"], ref["per_step"][i]))
        return {"latents": ref["per_step"][i].to(kwargs["latents"].dtype)}

    out, secs, _ = model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"],
                         num_inference_steps=n, guidance_scale=7.5, output_type="latent", callback_on_step_end=cb)
    assert model.scheduler.timesteps.tolist() == ref["timesteps"]
    floor = _torch_bf16_floor(world, Oc, kw, n, forced)
    floor_err = [_rel(f, r) for f, r in zip(floor, ref["per_step"])]
    print(f"\n[{name}] teacher-forced per-step relative max-abs: engine {max(errs):.3e}  torch-bf16 {max(floor_err):.3e}")
    assert max(errs) <= TF_TOL, errs
    assert max(errs) <= 1.5 * max(floor_err) + 2e-3, (errs, floor_err)


def test_pipeline_free_running_dpm(world):
    """configs[1] shape (DPM-Solver++ 2M, CFG 7.5) end to end at a reduced step count."""
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG

    P, Oc, kw, _ = _sched_pairs()["dpm++2"]
    n = 10
    ref = denoise(world["net"], Oc.from_config(SD15_SCHEDULER_CONFIG, **kw), world["pe"], world["ne"], world["lat"], n)
    model = world["make"]()
    model.scheduler = P.from_config(SD15_SCHEDULER_CONFIG, **kw)
    out, secs, x0s = model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"],
                           num_inference_steps=n, guidance_scale=7.5, output_type="latent")
    got = out.images.float()
    rng = ref["latents"].abs().max().item()
    err = (got - ref["latents"]).abs().max().item()
    print(f"\n[dpm++2 free-running {n} steps] max-abs {err:.3e} of range {rng:.2f}; loop {secs * 1e3:.1f} ms")
    assert len(x0s) == 0 or len(x0s) == n
    assert model.num_timesteps == n
    assert err <= FREE_TOL_REL * rng


def test_pipeline_lcm_no_cfg(world):
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG, LCMScheduler
    from sonicdiffusionbayeslab_b200 import schedulers as S

    n = 4
    go = torch.Generator(device="cuda").manual_seed(5)
    gp = torch.Generator(device="cuda").manual_seed(5)
    # noise must be drawn with the same torch call on both sides: bf16 latents
    lat16 = world["lat"].bfloat16()
    ref = denoise(world["net16"], LCMScheduler.from_config(SD15_SCHEDULER_CONFIG), world["pe"].bfloat16(),
                  world["ne"].bfloat16(), lat16, n, guidance_scale=0, generator=go)
    model = world["make"]()
    model.scheduler = S.LCMScheduler.from_config(SD15_SCHEDULER_CONFIG)
    out, _, x0s = model(prompt_embeds=world["pe"], latents=lat16, num_inference_steps=n, guidance_scale=0,
                        generator=gp, output_type="latent")
    assert model.scheduler.timesteps.tolist() == [999, 759, 499, 259]
    rng = ref["latents"].float().abs().max().item()
    err = (out.images.float() - ref["latents"].float()).abs().max().item()
    print(f"\n[lcm 4 steps, no CFG] engine vs torch-bf16 oracle: max-abs {err:.3e} of range {rng:.2f}")
    assert err <= FREE_TOL_REL * rng
    assert len(x0s) == 0      # output_type latent: x0 decodes are skipped


def test_deepcache_pattern_and_parity(world):
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG, PNDMScheduler
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    n, interval = 7, 3
    dc = DeepCacheOracle(world["net"])
    dc.set_params(cache_interval=interval, cache_branch_id=0)
    ref = denoise(world["net"], PNDMScheduler.from_config(SD15_SCHEDULER_CONFIG), world["pe"], world["ne"],
                  world["lat"], n, deepcache=dc)
    model = world["make"]()
    helper = DeepCacheSDHelper(pipe=model)
    helper.set_params(cache_interval=interval, cache_branch_id=0)
    helper.enable()
    errs = []

    def cb(pipe, i, t, kwargs):
        errs.append(_rel(kwargs["latents"], ref["per_step"][i]))
        return {"latents": ref["per_step"][i].to(kwargs["latents"].dtype)}

    out, _, _ = model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"],
                      num_inference_steps=n, guidance_scale=7.5, output_type="latent", callback_on_step_end=cb)
    helper.disable()
    # PNDM-7: timesteps 858,715,715,572,429,286,143,1 -> cur = 0,1,1,3,4,5,6,7 ; full iff cur % 3 == 0
    assert model.scheduler.timesteps.tolist() == ref["timesteps"]
    assert model.last_step_kinds == ["full", "cached", "cached", "full", "cached", "cached", "full", "cached"]
    print(f"\n[deepcache interval {interval}] teacher-forced per-step max-abs {max(errs):.3e}")
    assert max(errs) <= TF_TOL, errs


@pytest.mark.parametrize("branch", [1, 2, 3, 5])
def test_deepcache_other_branches(world, branch):
    """``cache_branch_id`` != 0 (DeepCacheSDHelper.set_params): the engine records a cached plan
    whose cut follows ``divmod(branch, 3)``; teacher-forced against the oracle's DeepCache restatement."""
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise
    from oracle.schedulers import SD15_SCHEDULER_CONFIG, DDIMScheduler
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    n, interval = 4, 2
    dc = DeepCacheOracle(world["net"])
    dc.set_params(cache_interval=interval, cache_branch_id=branch)
    ref = denoise(world["net"], DDIMScheduler.from_config(SD15_SCHEDULER_CONFIG), world["pe"], world["ne"],
                  world["lat"], n, deepcache=dc)
    model = world["make"]()
    model.scheduler = S.DDIMSchedulerMy.from_config(SD15_SCHEDULER_CONFIG)
    helper = DeepCacheSDHelper(pipe=model)
    helper.set_params(cache_interval=interval, cache_branch_id=branch)
    helper.enable()
    errs = []

    def cb(pipe, i, t, kwargs):
        errs.append(_rel(kwargs["latents"], ref["per_step"][i]))
        return {"latents": ref["per_step"][i].to(kwargs["latents"].dtype)}

    model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"], num_inference_steps=n,
          guidance_scale=7.5, output_type="latent", callback_on_step_end=cb)
    eng = model.engine(world["B"], True)                     # the engine recorded for THIS branch (helper still on)
    assert eng.cache_branch == branch
    helper.disable()
    assert model.last_step_kinds == ["full", "cached", "full", "cached"]
    full_n, full_f = eng.stats("full")
    cached_n, cached_f = eng.stats("cached")
    base_n, base_f = model.engine(world["B"], True).stats("cached")     # branch 0 (helper off -> default engine)
    assert base_n < cached_n < full_n and base_f < cached_f < full_f    # a deeper cut recomputes more
    print(f"\n[deepcache branch {branch}] teacher-forced per-step max-abs {max(errs):.3e}; launches full {full_n} cached {cached_n}")
    assert cached_n < full_n and max(errs) <= TF_TOL, errs


def test_two_schedulers_switch(world):
    from oracle import schedulers as O
    from oracle.pipeline import denoise_two
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    cfg = O.SD15_SCHEDULER_CONFIG
    ref = denoise_two(world["net"], O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg),
                      world["pe"], world["ne"], world["lat"], 10, 3)
    model = world["make"](M.StableDiffusionModelTwoSchedulers)
    model.scheduler_first = S.DDIMSchedulerMy.from_config(cfg)
    model.scheduler_second = S.DPMSolverScheduler.from_config(cfg)
    errs = []

    def cb(pipe, i, t, kwargs):
        errs.append(_rel(kwargs["latents"], ref["per_step"][i]))
        return {"latents": ref["per_step"][i].to(kwargs["latents"].dtype)}

    out, _, _ = model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"],
                      num_inference_steps_first=10, num_inference_steps_second=10, num_step_switch=3,
                      type_switch="closest", guidance_scale=7.5, output_type="latent", callback_on_step_end=cb)
    first, second = model.last_timesteps
    assert ([int(t) for t in first], [int(t) for t in second]) == ref["timesteps"]
    assert ref["timesteps"] == ([901, 801, 701], [701, 601, 501, 401, 301, 201, 101, 1])
    assert model.num_timesteps == 11
    print(f"\n[two schedulers 10/k=3] teacher-forced per-step max-abs {max(errs):.3e}")
    assert max(errs) <= TF_TOL, errs


def test_interleaved_schedulers(world):
    """models.py:733-1135: DPM-Solver++(2M) main grid of 10 steps, groups 1 and 3 replaced by one DDIM step each
    (DDIM set up on 10 // 2 = 5 steps, so its stride spans the group); history feed after every inter step."""
    from oracle import schedulers as O
    from oracle.pipeline import denoise_interleaved
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    cfg = O.SD15_SCHEDULER_CONFIG
    ref = denoise_interleaved(world["net"], O.DPMSolverScheduler.from_config(cfg), O.DDIMScheduler.from_config(cfg),
                              world["pe"], world["ne"], world["lat"], 10, [1, 3])
    assert ref["timesteps"] == ([901, 811, 721, 541, 451, 361, 181, 91], [721, 361])
    model = world["make"](M.StableDiffusionModelInterlivingSchedulers)
    model.scheduler_main = S.DPMSolverScheduler.from_config(cfg)
    model.scheduler_inter = S.DDIMSchedulerMy.from_config(cfg)
    errs = []

    def cb(pipe, i, t, kwargs):
        errs.append(_rel(kwargs["latents"], ref["per_step"][i]))
        return {"latents": ref["per_step"][i].to(kwargs["latents"].dtype)}

    model(prompt_embeds=world["pe"], negative_prompt_embeds=world["ne"], latents=world["lat"], num_inference_steps=10,
          interliving_steps=[1, 3], guidance_scale=7.5, output_type="latent", callback_on_step_end=cb)
    assert model.last_timesteps == ref["timesteps"]
    assert model.num_timesteps == 8
    print(f"\n[interleaved 10 / groups 1,3] teacher-forced per-step max-abs {max(errs):.3e}")
    assert len(errs) == 8 and max(errs) <= TF_TOL, errs


def test_cuda_graph_replay_equals_eager(world):
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    sd = dict(world["net"].state_dict())
    eng = UNetEngine(sd, n_latents=1, cfg_dup=True, io_dtype=torch.bfloat16, device=world["dev"])
    eng.x_in.copy_(world["lat"][:1].bfloat16())
    eng.set_context(torch.cat([world["ne"][:1], world["pe"][:1]]).bfloat16())
    eager = eng.forward(777.0).clone()
    eager_c = eng.forward(700.0, cached=True).clone()
    eng.capture_graphs()
    graph = eng.forward(777.0).clone()
    graph_c = eng.forward(700.0, cached=True).clone()
    torch.cuda.synchronize()
    assert torch.equal(eager, graph)
    assert torch.equal(eager_c, graph_c)


def test_main_py_runs_a_config_end_to_end(tmp_path):
    """``python main.py --config <yaml>`` -- the reference's entry point (/root/reference/main.py:10-24) -- drives the
    registry, the dpm_solver method, the engine, the VAE decode and the metric plugins and writes metrics.tsv."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = open(os.path.join(root, "tests", "golden", "e2e_smoke_config.yaml")).read()
    cfg = cfg.replace("./gpurun_out/e2e/", str(tmp_path) + "/")
    path = tmp_path / "cfg.yaml"
    path.write_text(cfg)
    r = subprocess.run([sys.executable, os.path.join(root, "main.py"), "--config", str(path)], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    tables = sorted(tmp_path.rglob("metrics.tsv"))
    assert len(tables) == 2                                        # one per sweep point (3 and 5 steps)
    rows = tables[-1].read_text().strip().splitlines()
    assert rows[0].split("\t")[:2] == ["nfe", "clip_score_gen_image"]
    assert [r.split("\t")[0] for r in rows[1:]] == ["3", "5"]      # nfe = number of UNet evaluations
    cols = rows[0].split("\t")
    assert all(float(r.split("\t")[cols.index("time_metric")]) > 0 for r in rows[1:])     # seconds per image, loop only
    assert all(r.split("\t")[cols.index("weights")] == "model:random-init,clip:random-init" for r in rows[1:])
    assert len(list(tmp_path.rglob("*.png"))) == 16                # 2 batches x 4 prompts x 2 sweep points


@pytest.mark.parametrize("n_img,latent", [(2, 32), (1, 64)])
def test_native_vae_decoder_matches_torch_module(cuda, n_img, latent):
    """VaeEngine (tcgen05 convs, fused GroupNorm, GEMM + row-softmax attention) against the oracle's PyTorch
    AutoencoderKL decoder (oracle/vae.py, fp32) on the same seeded weights; output is an image in roughly [-1, 1]."""
    from oracle.vae import make_vae
    from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine

    ref_mod = make_vae(29, dtype=torch.float32, device=cuda)
    sd = {k: v.detach() for k, v in ref_mod.state_dict().items()}
    eng = VaeEngine(sd, n_img=n_img, latent=latent, io_dtype=torch.float32, device=cuda)
    g = torch.Generator(device="cuda").manual_seed(5)
    z = torch.randn(n_img, 4, latent, latent, device=cuda, generator=g)
    out = eng.decode(z).clone()
    ref = ref_mod.decode(z)[0]
    bf = make_vae(29, dtype=torch.bfloat16, device=cuda).decode(z.bfloat16())[0].float()   # stock torch-bf16 error
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out - ref).abs().max().item() / scale
    err_bf = (bf - ref).abs().max().item() / scale
    assert out.shape == (n_img, 3, 8 * latent, 8 * latent)
    assert err < max(3e-2, 2.0 * err_bf), (err, err_bf)


def test_native_clip_towers_match_transformers(cuda):
    """ClipVisionEngine / ClipTextEngine (native LayerNorm, QKV / MLP GEMMs with QuickGELU epilogue, flash attention with
    the causal mask for text) against transformers.CLIPModel (fp32) on the same seeded ViT-B/16 weights: per-pair
    CLIP score 100*cos(f_img, f_txt) within the +-0.2 the north star allows."""
    from sonicdiffusionbayeslab_b200.clip_engine import ClipTextEngine, ClipVisionEngine
    from sonicdiffusionbayeslab_b200.metrics.metrics import make_clip_model

    from transformers import CLIPModel

    weights, tok = make_clip_model(None)
    model = CLIPModel(weights.config)
    model.load_state_dict(weights.state_dict())
    model = model.to(cuda).float().eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    n = 4
    g = torch.Generator(device="cuda").manual_seed(3)
    pixel = torch.randn(n, 3, 224, 224, device=cuda, generator=g)
    ids, mask = tok(["a photo of a cat", "two dogs running on the beach at sunset", "x", "a " * 60])
    ids, mask = ids.to(cuda), mask.to(cuda)
    with torch.no_grad():
        fi_ref = model.get_image_features(pixel_values=pixel)
        ft_ref = model.get_text_features(input_ids=ids, attention_mask=mask)
    fi_ref = getattr(fi_ref, "pooler_output", fi_ref)
    ft_ref = getattr(ft_ref, "pooler_output", ft_ref)
    vis = ClipVisionEngine(sd, n=n, device=cuda)
    txt = ClipTextEngine(sd, n=n, device=cuda)
    fi = vis.image_features(pixel).float()
    ft = txt.text_features(ids).float()
    torch.cuda.synchronize()

    def score(a, b):
        return 100 * torch.nn.functional.cosine_similarity(a, b, dim=-1)

    assert (fi - fi_ref).abs().max().item() < 3e-2 * fi_ref.abs().max().item()
    assert (ft - ft_ref).abs().max().item() < 3e-2 * ft_ref.abs().max().item()
    assert (score(fi, ft) - score(fi_ref, ft_ref)).abs().max().item() < 0.2


def test_native_prompt_encoder_matches_transformers(cuda):
    """encode_prompt (models.py:139-149) through ClipTextEngine (CLIP-L text tower, causal attention, final LayerNorm)
    against transformers.CLIPTextModel fp32 on the same seeded weights."""
    from transformers import CLIPTextModel

    from sonicdiffusionbayeslab_b200 import models as M

    with pytest.warns(RuntimeWarning, match="RANDOM-INIT"):
        model = M.StableDiffusionModel.from_pretrained("runwayml/stable-diffusion-v1-5",
                                                       torch_dtype=torch.bfloat16).to(cuda)
    assert model.weights_source == "random-init"
    prompts = ["a photo of an astronaut riding a horse", "", "sunset over mountains, oil painting"]
    got = model._encode(prompts).float()
    ref_mod = CLIPTextModel(model.text_encoder.config)
    ref_mod.load_state_dict(model.text_encoder.state_dict())
    ids, _ = model.tokenizer(prompts)
    with torch.no_grad():
        ref = ref_mod.to(cuda).float().eval()(ids.to(cuda))[0].float()
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (3, 77, 768)
    assert (got - ref).abs().max().item() < 3e-2 * ref.abs().max().item()


def _test_images(n=4, H=512, W=512):
    g = torch.Generator().manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    imgs = [torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)]
    for k in range(1, n):                                    # smooth, image-like content
        imgs.append(torch.stack([(127 + 120 * torch.sin(xx / (17.0 + 9 * k) + c) * torch.cos(yy / (11.0 + 5 * k)))
                                 .clamp(0, 255) for c in range(3)]).to(torch.uint8))
    return torch.stack(imgs)


@pytest.mark.parametrize("hw", [(512, 512), (480, 640)])
def test_clip_preprocess_kernel_matches_hf_processor(cuda, hw):
    """``sonic_clip_preprocess`` (one kernel: PIL-exact antialiased bicubic resize, centre crop, rescale, normalise)
    against transformers' CLIPImageProcessor: the PIL-backed one -- what the reference's pinned transformers 4.48 /
    torchmetrics CLIPScore run (/root/reference/src/metrics/metrics.py:25-41) -- to float rounding, the default
    (torchvision-backed) one to one uint8 level."""
    from transformers import CLIPImageProcessor

    from sonicdiffusionbayeslab_b200 import kernels as K

    imgs = _test_images(4, *hw)
    got = K.clip_preprocess(imgs.to(cuda)).cpu()
    assert got.shape == (4, 3, 224, 224)
    level = 1.0 / 255.0 / min(K.CLIP_STD)                     # one uint8 level in normalised units
    try:
        from transformers import CLIPImageProcessorPil

        want = CLIPImageProcessorPil()(images=[im for im in imgs], return_tensors="pt")["pixel_values"]
        assert (got - want).abs().max().item() <= 1e-6, (got - want).abs().max().item()
    except ImportError:
        pass
    # the default processor of transformers 5.x resizes with torchvision, which itself differs from PIL by up to two
    # uint8 levels on a fraction of a percent of the pixels
    want = CLIPImageProcessor()(images=[im for im in imgs], return_tensors="pt")["pixel_values"]
    d = (got - want).abs()
    assert d.max().item() <= 2.1 * level and (d > 0.5 * level).float().mean().item() < 0.02
    # fused quantise: float [0,1] images -> (x * 255).to(uint8) inside the kernel (base_experiment.py:198-199)
    f = torch.rand(2, 3, *hw, generator=torch.Generator().manual_seed(1))
    a = K.clip_preprocess(f.to(cuda))
    b = K.clip_preprocess((f * 255).to(torch.uint8).to(cuda))
    assert torch.equal(a, b)
    # fused patch cut: bf16 rows of the ViT patch-embedding GEMM
    patches = torch.zeros(4 * 196, 768, device=cuda, dtype=torch.bfloat16)
    K.clip_preprocess(imgs.to(cuda), patches_out=patches)
    ref = got.view(4, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(4 * 196, 768).bfloat16()
    assert torch.equal(patches.cpu(), ref)


def test_clip_score_end_to_end_matches_transformers(cuda):
    """uint8 512x512 images + prompts -> CLIP score through the product metric (native preprocess kernel + native
    towers) against transformers.CLIPModel fed by HF's CLIPImageProcessor (the torchmetrics CLIPScore recipe,
    SURVEY appendix A.5) on the same seeded weights: per-image and mean score within the north_star's +-0.2."""
    from transformers import CLIPImageProcessor, CLIPModel

    from sonicdiffusionbayeslab_b200.metrics.metrics import ClipScoreMetric

    with pytest.warns(RuntimeWarning, match="RANDOM-INIT"):
        metric = ClipScoreMetric().to(cuda)
    imgs = _test_images(4)
    text = ["a photo of a cat", "two dogs running on the beach at sunset", "x", "a " * 60]
    fi, ft = metric.features(imgs.to(cuda), text)
    got = 100 * (fi * ft).sum(-1)
    model = CLIPModel(metric.model.config)
    model.load_state_dict(metric.model.state_dict())
    model = model.to(cuda).float().eval()
    pixel = CLIPImageProcessor()(images=[im for im in imgs], return_tensors="pt")["pixel_values"].to(cuda)
    ids, mask = metric.tokenizer(text)
    with torch.no_grad():
        a = model.get_image_features(pixel_values=pixel)
        b = model.get_text_features(input_ids=ids.to(cuda), attention_mask=mask.to(cuda))
    a, b = getattr(a, "pooler_output", a), getattr(b, "pooler_output", b)
    want = 100 * torch.nn.functional.cosine_similarity(a, b, dim=-1)
    metric.update(imgs.to(cuda), text)
    mean_want = torch.clamp(want.mean(), min=0)
    print(f"\n[clip score e2e] per-image delta {(got - want).abs().max().item():.4f}, "
          f"mean {metric.compute().item():.4f} vs {mean_want.item():.4f}")
    assert (got - want).abs().max().item() < 0.2
    assert abs(metric.compute().item() - mean_want.item()) < 0.2


def test_calc_clip_score_script_sharded_equals_single(tmp_path):
    """``calc_clip_score.py`` (reference: /root/reference/calc_clip_score.py) on seeded synthetic images: one process
    vs two ranks with the feature all-gather (gloo here: two ranks share the one GPU)."""
    import os
    import re
    import socket
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(nproc):
        env = dict(os.environ, PYTHONWARNINGS="ignore")
        cmd = [sys.executable]
        if nproc > 1:
            with socket.socket() as s_:
                s_.bind(("127.0.0.1", 0))
                port = s_.getsockname()[1]
            env["SONIC_DIST_BACKEND"] = "gloo"
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                    "--master-port", str(port)]
        cmd += [os.path.join(root, "calc_clip_score.py"), "--synthetic", "10", "--batch_size", "4", "--gather_images"]
        r = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        return float(re.search(r"CLIP Score: ([-0-9.e]+)", r.stdout).group(1))

    one, two = run(1), run(2)
    assert abs(one - two) < 1e-4, (one, two)
