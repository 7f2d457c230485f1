"""The WIRING of the recorded launch plans against the oracle, on the CPU (tools/plan_interp.py).

An engine records every kernel launch once, with the operand pointers and the packed weights baked in.  Here the
engines are recorded over HOST buffers and the operator lists are executed by a small PyTorch interpreter that follows
the operator semantics documented in include/sonic.h (bf16 operands as stored, fp32 accumulation).  What comes out is
compared with the oracle networks -- so weight packing (tap-major 3x3, phase-form upsample, stride-2 views,
tile-interleaved GEGLU, LayerNorms folded into side chunks), concat order, skip connections, the time path and the
DeepCache cut of EVERY branch are checked without a GPU, at shapes the GPU parity tests do not run.  The kernels
themselves are what the GPU tests check; this checks that the plans ask them for the right computation.

Tolerance: the interpreted plan keeps activations in bf16 like the engine, the oracle is fp32: max-abs 1-1.4e-2 on an
eps of |max| ~1.5 (measured); a wiring error shows at 1e-1 ... 1.
"""
import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)

import plan_check  # noqa: E402
import plan_interp  # noqa: E402

TOL = 2.5e-2


@pytest.fixture(scope="module")
def unet():
    from oracle.unet import make_unet
    from sonicdiffusionbayeslab_b200.unet_engine import PackedWeights

    net = make_unet(29)
    return net, PackedWeights(dict(net.state_dict()), "cpu")


class _Session:
    """``with _Session() as s: eng = s.build(lambda: Engine(...)); s.run(eng.plan)``."""

    def __enter__(self):
        self._cm = plan_check.recording()
        self.tracker = self._cm.__enter__()
        self.interp = plan_interp.PlanInterpreter()
        record = self.tracker.on_op

        def on_op(name, args):
            self.interp.record(name, args)
            record(name, args)

        self.tracker.on_op = on_op
        return self

    def __exit__(self, *exc):
        return self._cm.__exit__(*exc)

    def build(self, make):
        self.tracker.raws.clear()
        return make()

    def run(self, plan):
        self.interp.run(plan.h)


def _inputs(n, hw, seed):
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randn(n, 77, 768, generator=g).bfloat16()
    x1 = torch.randn(n, 4, hw, hw, generator=g).bfloat16().float()
    x2 = torch.randn(n, 4, hw, hw, generator=g).bfloat16().float()
    return ctx, x1, x2


@pytest.mark.parametrize("n_latents,cfg,hw", [(1, True, 32), (3, False, 16)])
def test_full_unet_plan_computes_the_oracle_unet(unet, n_latents, cfg, hw):
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    net, packed = unet
    n = n_latents * (2 if cfg else 1)
    ctx, x, _ = _inputs(n, hw, seed=n)
    x = x[:n_latents]
    with _Session() as s:
        eng = s.build(lambda: UNetEngine(packed, n_latents=n_latents, cfg_dup=cfg, device="cpu", height=hw, width=hw))
        eng.ctx.copy_(ctx.reshape(n * 77, -1))
        eng.x_in.copy_(x)
        eng.t_dev.fill_(501.0)
        s.run(eng.plans["ctx"])
        s.run(eng.plans["full"])
        got = eng.eps.float().clone()
        assert s.tracker.problems == []
    with torch.no_grad():
        want = net(torch.cat([x] * 2) if cfg else x, torch.tensor(501), encoder_hidden_states=ctx.float())[0]
    err = (got - want).abs().max().item()
    print(f"\n[unet n_latents={n_latents} cfg={cfg} {hw}x{hw}] interpreted plan vs oracle: max-abs {err:.2e} "
          f"(|eps|max {want.abs().max():.2f})")
    assert want.abs().max() > 0.5 and err <= TOL, err


@pytest.mark.parametrize("branch", range(12))
def test_cached_plan_computes_the_deepcache_step_of_the_oracle(unet, branch):
    """A full step, then a cached step on different latents at a different timestep, for every ``cache_branch_id``
    (the GPU tests run branches 0, 1, 2, 3, 5): the cached plan must reproduce the oracle's DeepCache
    wrapper -- which is NOT what a full forward gives (printed beside it)."""
    from oracle.deepcache import DeepCacheOracle
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    net, packed = unet
    hw = 16
    ctx, x1, x2 = _inputs(2, hw, seed=100 + branch)
    x1, x2 = x1[:1], x2[:1]
    with _Session() as s:
        eng = s.build(lambda: UNetEngine(packed, n_latents=1, cfg_dup=True, device="cpu", height=hw, width=hw,
                                         cache_branch=branch))
        eng.ctx.copy_(ctx.reshape(2 * 77, -1))
        s.run(eng.plans["ctx"])
        eng.x_in.copy_(x1)
        eng.t_dev.fill_(801.0)
        s.run(eng.plans["full"])
        e1 = eng.eps.float().clone()
        eng.x_in.copy_(x2)
        eng.t_dev.fill_(301.0)
        s.run(eng.plans["cached"])
        e2 = eng.eps.float().clone()
        assert s.tracker.problems == []
    dc = DeepCacheOracle(net)
    dc.set_params(cache_interval=3, cache_branch_id=branch)
    with torch.no_grad():
        w1 = dc.forward(torch.cat([x1] * 2), torch.tensor(801), ctx.float(), 0)
        w2 = dc.forward(torch.cat([x2] * 2), torch.tensor(301), ctx.float(), 1)
        full2 = net(torch.cat([x2] * 2), torch.tensor(301), encoder_hidden_states=ctx.float())[0]
    err1, err2 = (e1 - w1).abs().max().item(), (e2 - w2).abs().max().item()
    gap = (w2 - full2).abs().max().item()
    print(f"\n[branch {branch}] full step {err1:.2e}, cached step {err2:.2e} vs the oracle; a full forward instead of "
          f"the cached step would differ by {gap:.2e}")
    assert err1 <= TOL and err2 <= TOL, (err1, err2)
    if branch <= 8:
        assert gap >= 2 * err2, (gap, err2)                      # the comparison tells the two apart


def test_vae_plan_computes_the_oracle_decoder():
    from oracle.vae import make_vae
    from sonicdiffusionbayeslab_b200.vae_engine import VaeEngine

    ref = make_vae(29, dtype=torch.float32)
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(5)).bfloat16().float()
    with _Session() as s:
        eng = s.build(lambda: VaeEngine(dict(ref.state_dict()), n_img=2, latent=16, io_dtype=torch.float32, device="cpu"))
        eng.z_in.copy_(z)
        s.run(eng.plan)
        got = eng.img.float().clone()
        assert s.tracker.problems == []
    with torch.no_grad():
        want = ref.decode(z)[0]
    scale = want.abs().max().item()
    err = (got - want).abs().max().item() / scale
    print(f"\n[vae 2 x 16x16 latents] interpreted plan vs oracle decoder: max-abs / range {err:.2e}")
    assert got.shape == (2, 3, 128, 128) and err <= 3e-2, err


def test_prompt_encoder_and_clip_tower_plans_compute_the_transformers_modules():
    from transformers import CLIPModel, CLIPTextModel

    from sonicdiffusionbayeslab_b200.clip_engine import ClipTextEngine, ClipVisionEngine
    from sonicdiffusionbayeslab_b200.metrics.metrics import make_clip_model
    from sonicdiffusionbayeslab_b200.text import make_text_encoder

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        text = make_text_encoder(29, None, dtype=torch.float32)
        clip, _ = make_clip_model(None)
    c = text.config
    hf_text = CLIPTextModel(c).eval()
    hf_text.load_state_dict(text.state_dict())
    ids = torch.randint(0, 49000, (2, 77), generator=torch.Generator().manual_seed(1))
    ids[:, 0], ids[:, -1] = 49406, 49407
    with _Session() as s:
        eng = s.build(lambda: ClipTextEngine(text.state_dict(), n=2, seq=c.max_position_embeddings, width=c.hidden_size,
                                             heads=c.num_attention_heads, layers=c.num_hidden_layers,
                                             mlp=c.intermediate_size, device="cpu"))
        emb = eng.tok[ids].float() + eng.pos.float()
        eng.x.copy_(emb.reshape(2 * 77, -1))
        s.run(eng.plan)
        got = eng.out.float().view(2, 77, -1).clone()
        assert s.tracker.problems == []
    with torch.no_grad():
        want = hf_text(input_ids=ids).last_hidden_state
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"\n[prompt encoder] interpreted plan vs transformers CLIPTextModel: max-abs / range {err:.2e}")
    assert err <= 3e-2, err

    hf = CLIPModel(clip.config).eval()
    hf.load_state_dict(clip.state_dict())
    px = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(2)).bfloat16().float()
    with _Session() as s:
        eng = s.build(lambda: ClipVisionEngine(clip.state_dict(), n=1, device="cpu"))
        n, g, p_ = 1, eng.grid, eng.patch
        eng.patches.copy_(px.view(n, 3, g, p_, g, p_).permute(0, 2, 4, 1, 3, 5).reshape(n * g * g, -1))
        eng.x.view(n, eng.seq, eng.width)[:, 0] = eng.cls_pos
        s.run(eng.pre_plan)
        s.run(eng.plan)
        got = eng.out.float().view(1, eng.seq, -1).clone()
        assert s.tracker.problems == []
    with torch.no_grad():
        want = hf.vision_model(pixel_values=px).last_hidden_state
    err = (got - want).abs().max().item() / want.abs().max().item()
    print(f"\n[clip image tower] interpreted plans vs transformers CLIPVisionModel: max-abs / range {err:.2e}")
    assert err <= 3e-2, err


@pytest.mark.parametrize("case", ["ddim_cfg", "dpmpp_deepcache_branch3", "lcm_no_cfg", "two_schedulers",
                                  "interleaved_rescale"])
def test_product_pipeline_over_interpreted_plans_equals_the_oracle_loop(unet, case, monkeypatch):
    """The whole product stack short of the kernels, on the CPU: ``StableDiffusionModel.__call__`` -> ``UNetEngine``
    (real plan recording, ``set_context`` / ``forward`` replaying the plans through the interpreter) -> fused scheduler
    step (float64 kernel model), against the oracle loop -- what ``__graft_entry__.smoke()`` checks on the GPU, plus a
    DeepCache run (cached plan of branch 3 every other step) and an LCM run without guidance."""
    from test_pipeline_host_cpu import _launch_in_place

    from oracle import schedulers as O
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise, denoise_interleaved, denoise_two
    from sonicdiffusionbayeslab_b200 import kernels as K
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200 import unet_engine as UE
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper
    from sonicdiffusionbayeslab_b200.text import HashTokenizer

    net, packed = unet
    hw, B = 16, 1
    g = torch.Generator().manual_seed(11)
    pe = torch.randn(B, 77, 768, generator=g).bfloat16().float()
    ne = torch.randn(B, 77, 768, generator=g).bfloat16().float()
    lat = torch.randn(B, 4, hw, hw, generator=g)
    cfg = O.SD15_SCHEDULER_CONFIG
    pp = dict(solver_order=2, algorithm_type="dpmsolver++")
    cls = M.StableDiffusionModel
    if case == "two_schedulers":                                    # DDIM for 2 steps, then DPM-Solver++ on the same grid
        cls, prod, orac, steps, guidance, dc = M.StableDiffusionModelTwoSchedulers, None, None, 5, 7.5, None
    elif case == "interleaved_rescale":                             # DPM-Solver++ main, DDIM on group 1, guidance_rescale
        cls, prod, orac, steps, guidance, dc = M.StableDiffusionModelInterlivingSchedulers, None, None, 6, 7.5, None
    elif case == "ddim_cfg":
        prod, orac, steps, guidance, dc = S.DDIMSchedulerMy.from_config(cfg), O.DDIMScheduler.from_config(cfg), 2, 7.5, None
    elif case == "lcm_no_cfg":
        prod, orac, steps, guidance, dc = S.LCMScheduler.from_config(cfg), O.LCMScheduler.from_config(cfg), 3, 0.0, None
    else:
        kw = dict(solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")
        prod, orac = S.DPMSolverScheduler.from_config(cfg, **kw), O.DPMSolverScheduler.from_config(cfg, **kw)
        steps, guidance, dc = 4, 7.5, (2, 3)

    with _Session() as s:
        monkeypatch.setattr(UE._Plan, "run", lambda plan, stream: s.interp.run(plan.h))
        monkeypatch.setattr(K, "stream_ptr", lambda: None)
        monkeypatch.setattr(S.FusedScheduler, "_launch", _launch_in_place)
        monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

        def engine(self, n_latents, cfg_dup):                       # _PipelineBase.engine without the CUDA requirement
            branch = self._deepcache["branch"] if self._deepcache else 0
            key = (n_latents, bool(cfg_dup), branch)
            if key not in self._engines:
                self._engines[key] = s.build(lambda: UE.UNetEngine(
                    packed, n_latents=n_latents, cfg_dup=cfg_dup, arch=self.arch, height=self.latent_size,
                    width=self.latent_size, io_dtype=self.dtype, device="cpu", cache_branch=branch))
            return self._engines[key]

        monkeypatch.setattr(M._PipelineBase, "engine", engine)
        model = cls(packed.sd, vae=None, text_encoder=None, tokenizer=HashTokenizer(),
                    scheduler=prod or S.PNDMScheduler.from_config(cfg), torch_dtype=torch.float32, latent_size=hw)
        helper = None
        if dc:
            helper = DeepCacheSDHelper(pipe=model)
            helper.set_params(cache_interval=dc[0], cache_branch_id=dc[1])
            helper.enable()
        kw = dict(generator=torch.Generator().manual_seed(3)) if case == "lcm_no_cfg" else {}
        common = dict(prompt_embeds=pe, negative_prompt_embeds=ne if guidance > 1 else None, latents=lat,
                      guidance_scale=guidance, output_type="latent")
        if case == "two_schedulers":
            model.scheduler_first = S.DDIMSchedulerMy.from_config(cfg)
            model.scheduler_second = S.DPMSolverScheduler.from_config(cfg, **pp)
            out, secs, _ = model(**common, num_inference_steps_first=steps, num_inference_steps_second=steps,
                                 num_step_switch=2, type_switch="closest")
        elif case == "interleaved_rescale":
            model.scheduler_main = S.DPMSolverScheduler.from_config(cfg, **pp)
            model.scheduler_inter = S.DDIMSchedulerMy.from_config(cfg)
            out, secs, _ = model(**common, num_inference_steps=steps, interliving_steps=[1], guidance_rescale=0.7)
        else:
            out, secs, _ = model(**common, num_inference_steps=steps, **kw)
        kinds = list(model.last_step_kinds)
        if helper:
            helper.disable()
        got = out.images.float().clone()
        assert s.tracker.problems == []
    oracle_dc = None
    if dc:
        oracle_dc = DeepCacheOracle(net)
        oracle_dc.set_params(cache_interval=dc[0], cache_branch_id=dc[1])
    kw = dict(generator=torch.Generator().manual_seed(3)) if case == "lcm_no_cfg" else {}
    if case == "two_schedulers":
        ref = denoise_two(net, O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg, **pp), pe, ne, lat,
                          steps, 2, "closest", guidance_scale=guidance)
    elif case == "interleaved_rescale":
        ref = denoise_interleaved(net, O.DPMSolverScheduler.from_config(cfg, **pp), O.DDIMScheduler.from_config(cfg), pe,
                                  ne, lat, steps, [1], guidance_scale=guidance, guidance_rescale=0.7)
    else:
        ref = denoise(net, orac, pe, ne, lat, steps, guidance_scale=guidance, deepcache=oracle_dc, **kw)
    rng = max(1.0, ref["latents"].abs().max().item())
    err = (got - ref["latents"]).abs().max().item() / rng
    print(f"\n[{case}] product pipeline over interpreted plans vs oracle loop: max-abs / range {err:.2e} "
          f"(range {rng:.2f}, steps {kinds})")
    if case == "two_schedulers":
        first, second = model.last_timesteps
        assert ([int(t) for t in first], [int(t) for t in second]) == ref["timesteps"]
    elif case == "interleaved_rescale":
        assert model.last_timesteps == ref["timesteps"]
    else:
        assert model.scheduler.timesteps.tolist() == ref["timesteps"]
        assert kinds == (["full", "cached", "full", "cached"] if dc else ["full"] * len(ref["timesteps"]))
    assert err <= 3e-2, err


@pytest.mark.parametrize("output_type", ["pil"])          # "np" / "pt": tests/test_host_cpu.py postprocess test
def test_decode_and_postprocess_over_interpreted_plans(unet, output_type, monkeypatch):
    """The tail of the call (models.py:287-335) with the real engines on the CPU: final latents and every step's
    ``x0_pred[0]`` through the VAE decoder plan (``VaeEngine``, interpreted), denormalise, ``postprocess`` to PIL /
    numpy, ``return_dict=False`` -- against the oracle loop + the oracle decoder."""
    import numpy as np
    from test_pipeline_host_cpu import _launch_in_place

    from oracle import schedulers as O
    from oracle.pipeline import denoise
    from oracle.vae import make_vae
    from sonicdiffusionbayeslab_b200 import kernels as K
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200 import unet_engine as UE
    from sonicdiffusionbayeslab_b200.text import HashTokenizer
    from sonicdiffusionbayeslab_b200.vae_spec import VaeWeights

    net, packed = unet
    vae = make_vae(29, dtype=torch.float32)
    hw, steps = 16, 2
    g = torch.Generator().manual_seed(21)
    pe = torch.randn(1, 77, 768, generator=g).bfloat16().float()
    ne = torch.randn(1, 77, 768, generator=g).bfloat16().float()
    lat = torch.randn(1, 4, hw, hw, generator=g)
    cfg = O.SD15_SCHEDULER_CONFIG
    with _Session() as s:
        monkeypatch.setattr(UE._Plan, "run", lambda plan, stream: s.interp.run(plan.h))
        monkeypatch.setattr(K, "stream_ptr", lambda: None)
        monkeypatch.setattr(S.FusedScheduler, "_launch", _launch_in_place)
        monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
        monkeypatch.setattr(M._PipelineBase, "engine", lambda self, n, dup: self._engines.setdefault(
            ("unet", n, dup), UE.UNetEngine(packed, n_latents=n, cfg_dup=dup, height=hw, width=hw, io_dtype=self.dtype,
                                            device="cpu")))
        monkeypatch.setattr(M._PipelineBase, "_decode",                      # the product's body minus the CUDA requirement
                            lambda self, z: self.vae_engine(z.shape[0]).decode(z.to(self.dtype)).clone())
        model = M.StableDiffusionModel(packed.sd, vae=VaeWeights(dict(vae.state_dict())), text_encoder=None,
                                       tokenizer=HashTokenizer(), scheduler=S.DDIMSchedulerMy.from_config(cfg),
                                       torch_dtype=torch.float32, latent_size=hw)
        (images, nsfw), secs, x0_images = model(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat,
                                                num_inference_steps=steps, guidance_scale=7.5, output_type=output_type,
                                                return_dict=False)
        assert s.tracker.problems == [] and nsfw is None
    ref = denoise(net, O.DDIMScheduler.from_config(cfg), pe, ne, lat, steps)
    sf = 0.18215

    def decoded(z):
        with torch.no_grad():
            return (vae.decode(z / sf)[0] / 2 + 0.5).clamp(0, 1)

    want = [decoded(ref["latents"])] + [decoded(x0) for x0 in ref["x0"]]
    got = [images] + list(x0_images)
    assert len(x0_images) == steps == len(ref["x0"])                          # one x0 preview per DDIM step (row 0 only)
    for a, b in zip(got, want):
        if output_type == "pil":
            assert isinstance(a, list) and a[0].size == (8 * hw, 8 * hw) and a[0].mode == "RGB"
            arr = np.stack([np.asarray(im) for im in a]).astype(np.float32) / 255.0
        else:
            assert isinstance(a, np.ndarray) and a.dtype == np.float32
            arr = a
        assert arr.shape == (1, 8 * hw, 8 * hw, 3)
        err = np.abs(arr - b.permute(0, 2, 3, 1).numpy()).max()
        assert err <= 4e-2, err                                              # [0, 1] images: bf16 decoder vs fp32 oracle
