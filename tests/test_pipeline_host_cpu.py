"""The product's pipeline ``call`` bodies (sonicdiffusionbayeslab_b200/models.py) end to end on the CPU, against the
per-step latents that EXECUTING THE REFERENCE'S OWN ``call`` bodies produced (tests/golden/reference_pins.npz,
``pipe/<case>/per_step``; /root/reference/src/models.py:21-335 / 338-730 / 733-1135 / 1138-1467 via oracle/refexec.py).

Only the two native pieces are replaced, by the checker's stand-ins: the UNet launch plan by the tiny oracle UNet the
fixtures were made with (a fake engine with the real engine's surface: ``x_in`` / ``set_context`` / ``forward`` /
``n_lat``) and the fused update kernel by its float64 model (tests/test_host_cpu.py ``_emulated_launch``).  Everything
else -- argument checks, timestep retrieval, classifier-free-guidance batching, scheduler dispatch, two-scheduler
switch, interleave partition and history feeding, skip mask, in-place latent update through ``out=``, RNG order,
callback contract, return arity -- is the product code that runs on the GPU.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import refpin_cases as RC  # noqa: E402
from test_host_cpu import _emulated_launch  # noqa: E402

PINS = np.load(os.path.join(HERE, "golden", "reference_pins.npz"))
META = json.load(open(os.path.join(HERE, "golden", "reference_pins.json")))


class FakeEngine:
    """The surface of ``UNetEngine`` the pipelines use, over the oracle UNet (fp32, CPU)."""

    def __init__(self, net, n_latents, cfg_dup):
        self.net, self.n_lat, self.cfg_dup = net, n_latents, cfg_dup
        self.x_in = torch.zeros(n_latents, RC.C, RC.HW, RC.HW)
        self.ctx = None
        self.calls = []

    def set_context(self, ctx):
        self.ctx = ctx.float()

    @torch.no_grad()
    def forward(self, t, cached=False):
        assert not cached
        self.calls.append(int(t))
        x = torch.cat([self.x_in] * 2) if self.cfg_dup else self.x_in
        return self.net(x, torch.tensor(int(t)), encoder_hidden_states=self.ctx)[0]


def _launch_in_place(self, coeffs, eps, eps_text, sample, hist=(), noise=None, want_m0=False, want_x0=True, out=None,
                     ring=True, post=None):
    """The kernel model, honouring ``out=`` like the kernel does (the pipelines update the resident latents in place)."""
    xn, m0, x0 = _emulated_launch(self, coeffs, eps, eps_text, sample, hist=hist, noise=noise, want_m0=want_m0,
                                  want_x0=want_x0, post=post)
    if out is not None:
        out.copy_(xn)
        xn = out
    return xn, m0, x0


@pytest.fixture(scope="module")
def net():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield RC.tiny_unet()
    torch.set_num_threads(n)


@pytest.fixture()
def harness(monkeypatch, net):
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer
    from sonicdiffusionbayeslab_b200.unet_engine import UNetArch

    monkeypatch.setattr(S.FusedScheduler, "_launch", _launch_in_place)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    engines = []

    def engine(self, n_latents, cfg_dup):
        engines.append(FakeEngine(net, n_latents, cfg_dup))
        return engines[-1]

    def encode_prompt(self, prompt, do_cfg, prompt_embeds=None, negative_prompt_embeds=None, negative_prompt=None):
        return prompt_embeds, negative_prompt_embeds             # fp32 embeddings as the fixtures used (no bf16 cast)

    monkeypatch.setattr(M._PipelineBase, "engine", engine)
    monkeypatch.setattr(M._PipelineBase, "encode_prompt", encode_prompt)

    def make(cls, scheduler):
        return cls({}, vae=None, text_encoder=None, tokenizer=HashTokenizer(), scheduler=scheduler, arch=UNetArch(),
                   torch_dtype=torch.float32, latent_size=RC.HW)

    return M, S, make, engines


def run_product_case(case, harness):
    """One pipeline case (tests/refpin_cases.py format) through the PRODUCT's ``call`` body over the fake engine.
    Returns dict(seen=[timesteps], per_step=[latents], shapes=[callback inputs], legacy=LegacyCallback or None, out,
    secs, x0, pipe, common, kw, cb, engine)."""
    M, S, make, engines = harness
    pe, ne, lat = RC.pipeline_inputs()
    default = S.PNDMScheduler.from_config(RC.SD15)
    seen, per_step = [], []

    edit, shapes = RC.context_edit(case), []

    def cb(pipe, i, t, kwargs):
        seen.append(int(t))
        per_step.append(kwargs["latents"].clone())
        shapes.append({k: list(v.shape) for k, v in kwargs.items()})
        new = edit(i, kwargs["prompt_embeds"]) if edit else None        # ctx_edit cases: replace the UNet context
        return {} if new is None else {"prompt_embeds": new}

    common = dict(prompt_embeds=pe, negative_prompt_embeds=ne, guidance_scale=case["guidance"], output_type="latent",
                  callback_on_step_end=cb)
    if edit:
        common["callback_on_step_end_tensor_inputs"] = list(RC.CB_INPUTS)
    legacy = RC.LegacyCallback() if case.get("legacy_cb") else None
    if legacy:
        common.update(callback=legacy, callback_steps=case["legacy_cb"]["callback_steps"])
    if case.get("gen_seed") is not None:
        common["generator"] = torch.Generator().manual_seed(case["gen_seed"])
    if case.get("rescale"):
        common["guidance_rescale"] = case["rescale"]                     # rescale_noise_cfg branch, models.py:244-250
    if case.get("n_img"):
        common["num_images_per_prompt"] = case["n_img"]                  # B x n latents, embeddings repeated in place
    if not case.get("draw_latents"):
        common["latents"] = lat
    else:
        common["height"], common["width"] = 8 * RC.HW, 8 * RC.HW
    kind = case["pipe"]
    if kind in ("single", "skip"):
        cls = M.StableDiffusionModel if kind == "single" else M.StableDiffusionModelSkipTimesteps
        pipe = make(cls, RC.make_scheduler(*case["sched"], module=S))
        kw = dict(num_inference_steps=case["steps"])
        if kind == "skip":
            kw["skip_timesteps"] = list(case["skip"])
    elif kind == "two":
        pipe = make(M.StableDiffusionModelTwoSchedulers, default)
        pipe.scheduler_first = RC.make_scheduler(*case["first"], module=S)
        pipe.scheduler_second = RC.make_scheduler(*case["second"], module=S)
        kw = dict(num_inference_steps_first=case["n1"], num_inference_steps_second=case["n1"],
                  num_step_switch=case["k"], type_switch=case["type_switch"])
    else:
        pipe = make(M.StableDiffusionModelInterlivingSchedulers, default)
        pipe.scheduler_main = RC.make_scheduler(*case["main"], module=S)
        pipe.scheduler_inter = RC.make_scheduler(*case["inter"], module=S)
        kw = dict(num_inference_steps=case["steps"], interliving_steps=list(case["groups"]))
    out, secs, x0 = pipe(**common, **kw)
    return dict(seen=seen, per_step=per_step, shapes=shapes, legacy=legacy, out=out, secs=secs, x0=x0, pipe=pipe,
                common=common, kw=kw, cb=cb, edit=edit, engine=engines[-1])


@pytest.mark.parametrize("name", list(RC.PIPELINE_CASES))
def test_product_call_bodies_reproduce_reference_source(name, harness):
    case = RC.PIPELINE_CASES[name]
    r = run_product_case(case, harness)
    seen, per_step, shapes, legacy, out, secs, x0, pipe = (r[k] for k in ("seen", "per_step", "shapes", "legacy", "out",
                                                                          "secs", "x0", "pipe"))
    common, kw, cb, edit = r["common"], r["kw"], r["cb"], r["edit"]
    want = torch.from_numpy(PINS[f"pipe/{name}/per_step"])
    assert seen == META["pipeline_timesteps"][name]                      # integer schedule: bit-exact
    assert r["engine"].calls == seen                                     # one UNet evaluation per executed step
    assert len(per_step) == want.shape[0]
    scale = max(1.0, want.abs().max().item())
    worst = max((g - w).abs().max().item() for g, w in zip(per_step, want)) / scale
    print(f"\n[{name}] product call body vs reference source: worst per-step max-abs / range {worst:.2e}")
    assert worst <= 5e-6, (name, worst)                                  # fp32 host path, float64 kernel model
    assert torch.equal(out.images, per_step[-1]) and secs >= 0
    if edit:                                                             # what the callback is handed, models.py:263-267
        assert shapes == META["callback_shapes"][name]
    else:
        assert all(list(d) == ["latents"] for d in shapes)
    if legacy:                                                           # deprecated callback, models.py:275-282
        want_calls = META["legacy_callback"][name]
        assert [c[:2] for c in legacy.calls] == [c[:2] for c in want_calls]
        assert all(abs(a[2] - b[2]) <= 1e-5 * abs(b[2]) for a, b in zip(legacy.calls, want_calls))
    info = META["pipeline_info"][name]
    assert pipe.num_timesteps == info["num_timesteps"]
    assert x0 == []                                                      # output_type="latent": nothing is decoded
    common.pop("callback", None)
    tup = pipe(**{**common, "callback_on_step_end": cb if edit else None,     # the edit is part of the computation
                  **({"generator": torch.Generator().manual_seed(case["gen_seed"])} if case.get("gen_seed") is not None
                     else {})}, **kw, return_dict=False)
    assert isinstance(tup[0], tuple) and tup[0][1] is None and torch.equal(tup[0][0], out.images)   # models.py:322-323


class DeepCacheFakeEngine(FakeEngine):
    """Fake engine whose cached plan is the oracle's DeepCache wrapper: the PRODUCT decides
    which steps replay the cached plan; the oracle's own full / skip rule must agree on every call."""

    def __init__(self, net, n_latents, cfg_dup, t_list, interval, branch):
        from oracle.deepcache import DeepCacheOracle

        super().__init__(net, n_latents, cfg_dup)
        self.t_list, self.flags = t_list, []
        self.dc = DeepCacheOracle(net)
        self.dc.set_params(cache_interval=interval, cache_branch_id=branch)

    @torch.no_grad()
    def forward(self, t, cached=False):
        cur = self.t_list.index(int(t))
        start = cur if self.dc.start is None else self.dc.start
        assert ((cur - start) % self.dc.interval != 0) == cached, (int(t), cur, cached)
        self.calls.append(int(t))
        self.flags.append(bool(cached))
        x = torch.cat([self.x_in] * 2) if self.cfg_dup else self.x_in
        return self.dc.forward(x, torch.tensor(int(t)), self.ctx, cur)


def _random_deepcache_cases(n=14, seed=77):
    import random

    rng = random.Random(seed)
    return [(rng.choice(["ddim", "pndm", "dpm"]), rng.randint(3, 16), rng.randint(2, 6), rng.randint(0, 11))
            for _ in range(n)]


@pytest.mark.parametrize("sched,steps,interval,branch", [("ddim", 12, 3, 0), ("pndm", 7, 2, 0), ("pndm", 9, 5, 4),
                                                          ("ddim", 10, 4, 7)] + _random_deepcache_cases())
def test_product_deepcache_loop_equals_oracle(sched, steps, interval, branch, harness, net, monkeypatch):
    """``DeepCacheSDHelper`` + the pipeline's full / cached decision (deep_cache.py:24-29,58; appendix A.4 -- with
    PLMS the repeated timestep maps to its FIRST index) against the oracle's wrapper semantics, through the product's
    ``call``: same per-step latents, the cached plan replayed on exactly the steps the oracle skips."""
    from oracle import schedulers as O
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise
    from sonicdiffusionbayeslab_b200.deepcache import DeepCacheSDHelper

    M, S, make, engines = harness
    pe, ne, lat = RC.pipeline_inputs()
    pipe = make(M.StableDiffusionModel, RC.make_scheduler(sched, {}, module=S))

    def engine(self, n_latents, cfg_dup):
        assert self._deepcache["branch"] == branch
        engines.append(DeepCacheFakeEngine(net, n_latents, cfg_dup, list(self.scheduler._timesteps_host), interval,
                                           branch))
        return engines[-1]

    monkeypatch.setattr(M._PipelineBase, "engine", engine)
    helper = DeepCacheSDHelper(pipe=pipe)
    helper.set_params(cache_interval=interval, cache_branch_id=branch)
    helper.enable()
    per_step = []
    out, _, _ = pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, num_inference_steps=steps,
                     guidance_scale=7.5, output_type="latent",
                     callback_on_step_end=lambda p, i, t, kw: per_step.append(kw["latents"].clone()) or {})
    helper.disable()
    dc = DeepCacheOracle(net)
    dc.set_params(cache_interval=interval, cache_branch_id=branch)
    ref = denoise(net, RC.make_scheduler(sched, {}, module=O), pe, ne, lat, steps, deepcache=dc)
    eng = engines[-1]
    assert eng.calls == ref["timesteps"] and pipe.last_step_kinds == ["cached" if f else "full" for f in eng.flags]
    assert eng.flags[0] is False and any(eng.flags)
    worst = max((g - w).abs().max().item() for g, w in zip(per_step, ref["per_step"]))
    assert worst <= 5e-6 * max(1.0, ref["latents"].abs().max().item()), worst
    assert pipe._deepcache is None


@pytest.mark.parametrize("sched,over,steps", [
    ("dpm", dict(solver_order=2, algorithm_type="dpmsolver++", thresholding=True, sample_max_value=2.5), 8),
    ("dpm", dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min", prediction_type="v_prediction",
                 thresholding=True, sample_max_value=3.0), 8),
    ("ddim", dict(prediction_type="v_prediction", clip_sample=True, clip_sample_range=1.5), 6),
    ("lcm", dict(prediction_type="sample"), 4),
    ("pndm", dict(prediction_type="v_prediction"), 6),
])
def test_product_loop_with_prediction_types_and_x0_postprocessing(sched, over, steps, harness, net):
    """The scheduler options added on top of the SD-v1.5 defaults (every ``prediction_type``, ``thresholding``,
    ``clip_sample``) through the product's classifier-free-guidance loop -- the fused CFG + update (+ quantile) path --
    against the oracle loop with the oracle schedulers."""
    from oracle import schedulers as O
    from oracle.pipeline import denoise

    M, S, make, engines = harness
    pe, ne, lat = RC.pipeline_inputs()
    guidance = 0.0 if sched == "lcm" else 7.5
    gens = [torch.Generator().manual_seed(5) for _ in range(2)]
    pipe = make(M.StableDiffusionModel, RC.make_scheduler(sched, over, module=S))
    per_step = []
    pipe(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, num_inference_steps=steps, guidance_scale=guidance,
         generator=gens[0], output_type="latent",
         callback_on_step_end=lambda p, i, t, kw: per_step.append(kw["latents"].clone()) or {})
    ref = denoise(net, RC.make_scheduler(sched, over, module=O), pe, ne, lat, steps, guidance_scale=guidance,
                  generator=gens[1])
    assert engines[-1].calls == ref["timesteps"]
    worst = max((g - w).abs().max().item() for g, w in zip(per_step, ref["per_step"]))
    assert worst <= 2e-5 * max(1.0, max(w.abs().max().item() for w in ref["per_step"])), worst


@pytest.mark.parametrize("n,skip", [(20, [2, 3, 9, 15, 16]), (10, [0, 4, 9])])
def test_skip_steps_gpu_harness_dry_run(n, skip, harness, net):
    """tests/test_skip_steps_gpu.py's harness (tests/skip_case.py) over the fake engine: the loop-index / executed-step
    mapping, the forcing and the structural assertions of the GPU test hold on the CPU, where product and oracle differ
    only by the float64 kernel model (also with the first and the last index skipped)."""
    import skip_case

    M, S, make, engines = harness
    pe, ne, lat = RC.pipeline_inputs()
    noise = torch.randn(lat.shape, generator=torch.Generator().manual_seed(3))
    model = make(M.StableDiffusionModelSkipTimesteps, S.PNDMScheduler.from_config(RC.SD15))
    r = skip_case.run(model, net, None, pe, ne, lat, noise, n, skip)
    executed = [i for i in range(n) if i not in skip]
    assert [i for i, _ in r["seen"]] == executed
    assert [t for _, t in r["seen"]] == [r["timesteps"][i] for i in executed] == engines[-1].calls
    assert model.num_timesteps == n and len(model.last_step_kinds) == len(executed)
    assert r["x0"] == [] and r["torch_bf16"] is None
    assert max(r["engine"]) <= 5e-6 * max(1.0, r["xmax"]), r["engine"]
