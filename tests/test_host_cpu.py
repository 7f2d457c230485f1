"""CPU tests of the product's host logic: C-ABI surface, registries, scheduler schedules and the
coefficient reduction of every scheduler (checked by emulating the fused kernel in float64)."""
import ctypes
import dataclasses
import json
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "schedule_kat.json")))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sonic.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sonic_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from sonicdiffusionbayeslab_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libsonic.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sonic.h but not exported"
    lib.sonic_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sonic_version()


def test_library_is_blackwell_native():
    """SASS must contain tcgen05 MMA, TMEM loads and TMA loads (B200_PROFILING.md evidence table)."""
    from sonicdiffusionbayeslab_b200 import _lib

    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    # the CTA-pair GEMM (cta_group::2 MMA, multicast commit, pair TMA loads), cluster barriers (pair GEMM, GroupNorm)
    # and the packed fp32 pair arithmetic of the softmax / GEMM epilogues
    for mnemonic in ("UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMALDG.4D.2CTA", "UCGABAR_WAIT", "FFMA2", "FADD2", "FMUL2"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass            # no legacy mma.sync tensor path
    # programmatic dependent launch (SONIC_PDL): every kernel that can be launched with the attribute -- 2 GEMM,
    # 16 attention instantiations, cluster GroupNorm, ln_side -- waits (griddepcontrol.wait -> ACQBULK) exactly once,
    # and with the stock trigger policy signals (launch_dependents -> PREEXIT) exactly once
    assert sass.count("ACQBULK") == sass.count("PREEXIT") == 20


def test_missing_library_fails_loudly(monkeypatch):
    from sonicdiffusionbayeslab_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libsonic.so")
    with pytest.raises(_lib.SonicError):
        _lib.lib()


def test_registries_hold_reference_names():
    import sonicdiffusionbayeslab_b200 as pkg  # noqa: F401  (registration is an import side effect)
    from sonicdiffusionbayeslab_b200.registry import (methods_registry, metrics_registry, models_registry,
                                                      schedulers_registry)

    for n in ("stable_diffusion_model", "stable_diffusion_model_two_schedulers",
              "stable_diffusion_model_interliving_schedulers", "stable_diffusion_model_skip_timesteps"):
        assert n in models_registry
    for n in ("dpm_solver_scheduler", "ddim_scheduler", "lcm_scheduler"):
        assert n in schedulers_registry
    for n in ("ddim", "dpm_solver", "deep_cache", "consistency_model", "two_schedulers", "default"):
        assert n in methods_registry
    for n in ("clip_score", "time_metric", "image_reward", "fid"):
        assert n in metrics_registry
    fields = [f.name for f in dataclasses.fields(schedulers_registry.args["dpm_solver_scheduler"])]
    assert "solver_order" in fields and "algorithm_type" in fields


def test_class_registry_contract():
    from sonicdiffusionbayeslab_b200.utils.class_registry import MISSING, ClassRegistry

    reg = ClassRegistry()

    @reg.add_to_registry("thing")
    class Thing:
        def __init__(self, a, b=None, c=3, *args, **kwargs):
            pass

    assert reg["thing"] is Thing
    fs = {f.name: f for f in dataclasses.fields(reg.args["thing"])}
    assert set(fs) == {"a", "b", "c"}
    assert fs["a"].default == MISSING and fs["b"].default is None and fs["c"].default == 3


def _pairs():
    from oracle import schedulers as O
    from sonicdiffusionbayeslab_b200 import schedulers as S

    return [
        (S.DDIMSchedulerMy, O.DDIMScheduler, {}, 20, {}),
        (S.DDIMSchedulerMy, O.DDIMScheduler, {}, 7, {"eta": 0.5}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=2, algorithm_type="dpmsolver++"), 25, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=2, algorithm_type="dpmsolver++"), 6, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 10, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=2, algorithm_type="dpmsolver", final_sigmas_type="sigma_min", solver_type="heun"), 10, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=3, algorithm_type="dpmsolver++"), 20, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=3, algorithm_type="dpmsolver", final_sigmas_type="sigma_min"), 16, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=1, algorithm_type="dpmsolver++"), 5, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=2, algorithm_type="sde-dpmsolver++"), 8, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=2, algorithm_type="sde-dpmsolver", final_sigmas_type="sigma_min"), 8, {}),
        (S.LCMScheduler, O.LCMScheduler, {}, 4, {}),
        (S.PNDMScheduler, O.PNDMScheduler, {}, 12, {}),
        (S.DDIMSchedulerMy, O.DDIMScheduler, dict(prediction_type="v_prediction"), 10, {}),
        (S.DDIMSchedulerMy, O.DDIMScheduler, dict(prediction_type="sample"), 10, {"eta": 0.3}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=2, algorithm_type="dpmsolver++",
                                                          prediction_type="v_prediction"), 12, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=3, algorithm_type="dpmsolver", final_sigmas_type="sigma_min", prediction_type="sample"), 12, {}),
        (S.LCMScheduler, O.LCMScheduler, dict(prediction_type="v_prediction"), 4, {}),
        (S.LCMScheduler, O.LCMScheduler, dict(prediction_type="sample"), 3, {}),
        (S.PNDMScheduler, O.PNDMScheduler, dict(prediction_type="v_prediction"), 10, {}),
        (S.DDIMSchedulerMy, O.DDIMScheduler, dict(clip_sample=True, clip_sample_range=1.5), 10, {}),
        (S.DDIMSchedulerMy, O.DDIMScheduler, dict(thresholding=True, sample_max_value=2.5, prediction_type="v_prediction"),
         10, {"eta": 0.2}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler, dict(solver_order=2, algorithm_type="dpmsolver++", thresholding=True,
                                                          sample_max_value=3.0), 10, {}),
        (S.DPMSolverScheduler, O.DPMSolverScheduler,
         dict(solver_order=2, algorithm_type="sde-dpmsolver", final_sigmas_type="sigma_min", thresholding=True,
              dynamic_thresholding_ratio=0.9, sample_max_value=2.0), 8, {}),
    ]


def _emulated_post(x0, x, post):
    """sonic_latent_update_post (include/sonic.h): clip or dynamic thresholding of x0, then m0 = p_x x + p_0 x0'."""
    if post["mode"] == 2:
        b = x0.shape[0]
        s_ = torch.quantile(x0.reshape(b, -1).abs().float(), post["ratio"], dim=1).clamp(min=1, max=post["max_value"])
        s_ = s_.to(x0.dtype).reshape((b,) + (1,) * (x0.dim() - 1))
        x0 = torch.maximum(torch.minimum(x0, s_), -s_) / s_
    else:
        x0 = x0.clamp(-post["clip"], post["clip"])
    return x0, post["p_x"] * x + post["p_0"] * x0


def _emulated_launch(self, coeffs, eps, eps_text, sample, hist=(), noise=None, want_m0=False, want_x0=True, out=None,
                     ring=True, post=None):
    """float64 model of sonic_latent_update (include/sonic.h) for fp32 tensors on the CPU."""
    c = {k: float(coeffs.get(k, 0.0)) for k in ("guidance", "m_x", "m_e", "x0_x", "x0_e", "c_x", "c_e", "c_m0",
                                                "c_h1", "c_h2", "c_h3", "c_z")}
    d = torch.float64
    e = eps.to(d) if eps_text is None else eps.to(d) + c["guidance"] * (eps_text.to(d) - eps.to(d))
    x = sample.to(d)
    m0 = c["m_x"] * x + c["m_e"] * e
    x0 = c["x0_x"] * x + c["x0_e"] * e
    if post is not None:
        x0, m0 = _emulated_post(x0, x, post)
    h = [t.to(d) for t in hist] + [torch.zeros_like(x)] * (3 - len(hist))
    z = torch.zeros_like(x) if noise is None else noise.to(d)
    xn = c["c_x"] * x + c["c_e"] * e + c["c_m0"] * m0 + c["c_h1"] * h[0] + c["c_h2"] * h[1] + c["c_h3"] * h[2] + c["c_z"] * z
    f = sample.dtype
    return xn.to(f), (m0.to(f) if want_m0 else None), (x0.to(f) if want_x0 else None)


@pytest.mark.parametrize("idx", range(24))
def test_scheduler_coefficients_reproduce_oracle_updates(idx, monkeypatch):
    """Every scheduler's reduction to linear-combination coefficients, step by step, against the
    oracle's literal formulas (fp32, CPU, no GPU needed)."""
    from oracle.schedulers import SD15_SCHEDULER_CONFIG
    from sonicdiffusionbayeslab_b200 import schedulers as S

    monkeypatch.setattr(S.FusedScheduler, "_launch", _emulated_launch)
    P, Oc, kw, n, step_kw = _pairs()[idx]
    ps, os_ = P.from_config(SD15_SCHEDULER_CONFIG, **kw), Oc.from_config(SD15_SCHEDULER_CONFIG, **kw)
    ps.set_timesteps(n)
    os_.set_timesteps(n)
    assert ps.timesteps.tolist() == os_.timesteps.tolist()
    assert torch.equal(ps.alphas_cumprod, os_.alphas_cumprod)
    if hasattr(os_, "sigmas"):
        assert torch.equal(ps.sigmas, os_.sigmas)
    g = torch.Generator().manual_seed(idx)
    xp = xo = torch.randn(2, 4, 8, 8, generator=g)
    needs_gen = "generator" in step_kw or P is S.LCMScheduler or "sde" in kw.get("algorithm_type", "") or step_kw.get("eta")
    gp, go = torch.Generator().manual_seed(99), torch.Generator().manual_seed(99)
    for t in os_.timesteps:
        eps = torch.randn(2, 4, 8, 8, generator=g)
        kp, ko = dict(step_kw), dict(step_kw)
        if needs_gen:
            kp["generator"], ko["generator"] = gp, go
        rp, ro = ps.step(eps, t, xp, **kp), os_.step(eps, t, xo, **ko)
        assert len(rp) == len(ro)
        for a, b in zip(rp, ro):
            scale = max(1.0, b.abs().max().item())
            assert (a - b).abs().max().item() / scale < 2e-5, (type(ps).__name__, kw, int(t))
        xp, xo = rp[0], ro[0]
    assert ps.step_index == os_.step_index or P is S.DDIMSchedulerMy or P is S.PNDMScheduler


def test_product_schedules_match_golden():
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    cfg = M.SD15_SCHEDULER_CONFIG
    d = S.DDIMSchedulerMy.from_config(cfg)
    d.set_timesteps(20)
    assert d.timesteps.tolist() == KAT["ddim_20"]
    p = S.DPMSolverScheduler.from_config(cfg, solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")
    p.set_timesteps(25)
    assert p.timesteps.tolist() == KAT["dpm_25"]
    l = S.LCMScheduler.from_config(cfg)
    l.set_timesteps(4)
    assert l.timesteps.tolist() == KAT["lcm_4"]
    q = S.PNDMScheduler.from_config(cfg)
    q.set_timesteps(50)
    assert q.timesteps.tolist() == KAT["pndm_50"]
    # typo'd / unknown keys are dropped silently, as diffusers' from_config does (two_schedulers.py:51)
    S.DPMSolverScheduler.from_config(cfg, sovler_order=3)
    with pytest.raises(ValueError):
        S.DPMSolverScheduler.from_config(cfg, algorithm_type="dpmsolver", final_sigmas_type="zero")


def test_two_scheduler_switch_host_logic():
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    cfg = M.SD15_SCHEDULER_CONFIG
    a, b = S.DDIMSchedulerMy.from_config(cfg), S.DPMSolverScheduler.from_config(cfg)
    a.set_timesteps(10)
    b.set_timesteps(timesteps=a.timesteps.numpy())
    sw = M.StableDiffusionModelTwoSchedulers.switch_timestamp
    for mode in ("closest", "left_closest", "right_closest"):
        first, second = sw(None, a.timesteps, b.timesteps, 3, mode)
        assert ([int(t) for t in first], [int(t) for t in second]) == tuple(KAT["two_10_3"])


def test_interleaved_host_logic_and_config():
    """Interleaved-scheduler partition (src/models.py:944-961) equals the oracle's; the example YAML parses and its
    model / method / scheduler names resolve through the registries."""
    import os

    from oracle.pipeline import interleave_partition
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.config import load as load_config
    from sonicdiffusionbayeslab_b200.registry import methods_registry, models_registry, schedulers_registry

    main = S.DPMSolverScheduler.from_config(M.SD15_SCHEDULER_CONFIG)
    main.set_timesteps(20)
    ts = main.timesteps.tolist()
    part = M.StableDiffusionModelInterlivingSchedulers.partition
    for order, groups in ((2, [2, 5]), (2, [0, 9]), (3, [1, 4]), (1, [3]), (2, [])):
        assert part(ts, order, groups) == interleave_partition(ts, order, groups)
    kept, inter = part(ts, 2, [2, 5])
    assert len(kept) == 18 and inter == [ts[4], ts[10]] and ts[5] not in kept and ts[11] not in kept

    cfg = load_config(os.path.join(ROOT, "configs", "interliving_schedulers_config.yaml"))
    assert models_registry[cfg.model.model_name] is M.StableDiffusionModelInterlivingSchedulers
    assert methods_registry[cfg.experiment.method].__name__ == "InterlivingSchedulersMethod"
    assert schedulers_registry[cfg.scheduler.scheduler_main] is S.DPMSolverScheduler
    assert schedulers_registry[cfg.scheduler.scheduler_inter] is S.DDIMSchedulerMy
    assert [list(v) for v in cfg.experiment_params.interliving_steps] == [[2, 5], [1, 3, 5, 7]]
    with pytest.raises(ValueError):                    # the reference indexes sigmas[None] here (C-4 analogue)
        main.feed_history(None, None, 0.0, None)


def test_unet_spec_matches_oracle_keys():
    from oracle.unet import UNet2DConditionModel
    from sonicdiffusionbayeslab_b200.unet_spec import unet_param_shapes

    with torch.device("meta"):
        net = UNet2DConditionModel()
    want = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    got = dict(unet_param_shapes())
    assert got == want
    assert sum(torch.Size(s).numel() for s in got.values()) == KAT["unet_params"]


def test_step_requires_cuda():
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    s = S.DDIMSchedulerMy.from_config(M.SD15_SCHEDULER_CONFIG)
    s.set_timesteps(4)
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.step(x, s.timesteps[0], x)


@pytest.mark.parametrize("branch", list(range(12)))
def test_deepcache_branch_layer_sets_match_the_oracle(branch):
    """The engine's cached-plan walk (which layers a non-refresh step recomputes, which feature it reads) against
    the oracle's restatement of ``DeepCacheSDHelper.is_skip_step`` executed on a tiny UNet, for every branch."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    import refpin_cases as RC
    from oracle.deepcache import DeepCacheOracle
    from sonicdiffusionbayeslab_b200.unet_engine import deepcache_runs

    net = RC.tiny_unet()
    dc = DeepCacheOracle(net)
    dc.set_params(cache_interval=5, cache_branch_id=branch)
    pe, ne, lat = RC.pipeline_inputs()
    x = torch.cat([lat, lat])
    ctx = torch.cat([ne, pe])
    ran = []
    orig = dc._wrap

    def spy(key, block_i, layer_i, blocktype, fn):
        def fn2():
            ran.append(key)
            return fn()
        return orig(key, block_i, layer_i, blocktype, fn2)

    dc._wrap = spy
    with torch.no_grad():
        dc.forward(x, torch.tensor(500), ctx, 0)          # refresh step
        ran.clear()
        dc.forward(x, torch.tensor(400), ctx, 1)          # cached step
    nb, L = 4, 2
    down_runs, up_runs, first_up = deepcache_runs(branch, nb, L)
    want_down = {(b, j) for b in range(nb) for j in range(L) if down_runs(b, j)}
    want_ds = {b for b in range(nb - 1) if down_runs(b, L)}
    want_up = {(b, j) for b in range(nb) for j in range(L + 1) if up_runs(b, j)}
    got_down = {(k[2], k[3]) for k in ran if k[:2] == ("down", "resnet")}
    got_ds = {k[2] for k in ran if k[:2] == ("down", "downsampler")}
    got_up = {(nb - 1 - k[2], L - k[3]) for k in ran if k[:2] == ("up", "resnet")}       # reversed -> forward indices
    assert (got_down, got_ds, got_up) == (want_down, want_ds, want_up)
    assert not any(k[0] == "mid" for k in ran)
    assert min(want_up) == first_up
    got_us = {nb - 1 - k[2] for k in ran if k[:2] == ("up", "upsampler")}
    assert got_us == {b for b in range(nb - 1) if any(up_runs(b, j) for j in range(L + 1))}


def test_scheduler_swap_keeps_hidden_config_keys():
    """``Other.from_config(pipe.scheduler.config)`` (base_experiment.py:69-72): keys a scheduler does not take
    survive in its config, so DDIM built from the stock PNDM scheduler still sees ``clip_sample=False``."""
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    pndm = S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG)
    assert "clip_sample" not in S.PNDMScheduler._defaults and pndm.config.clip_sample is False
    ddim = S.DDIMSchedulerMy.from_config(pndm.config)
    assert ddim.config.clip_sample is False and ddim.config.steps_offset == 1
    lcm = S.LCMScheduler.from_config(S.DPMSolverScheduler.from_config(pndm.config, solver_order=3).config)
    assert lcm.config.clip_sample is False and lcm.config.solver_order == 3


@pytest.mark.parametrize("hw", [(512, 512), (480, 640), (224, 224)])
def test_pil_resample_tables_are_bit_exact(hw):
    """The coefficient tables the native CLIP-preprocess kernel consumes (kernels.pil_bicubic_coeffs: Pillow's
    precompute_coeffs + normalize_coeffs_8bpc restated) and the kernel's two-pass integer arithmetic, emulated in
    numpy, against PIL itself -- bit-exact uint8 images."""
    import numpy as np
    from PIL import Image

    from sonicdiffusionbayeslab_b200.kernels import pil_bicubic_coeffs

    H, W = hw
    short, long_ = min(H, W), max(H, W)
    nh, nw = (224, int(224 * long_ / short)) if H <= W else (int(224 * long_ / short), 224)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BICUBIC))
    hb, hk, _ = pil_bicubic_coeffs(W, nw)
    vb, vk, _ = pil_bicubic_coeffs(H, nh)
    got = np.zeros((nh, nw, 3), np.uint8)
    for c in range(3):
        tmp = np.zeros((H, nw), np.uint8)
        for ox in range(nw):
            xmin, cnt = hb[ox]
            ss = (1 << 21) + (img[:, xmin:xmin + cnt, c].astype(np.int64) * np.array(hk[ox][:cnt], np.int64)).sum(1)
            tmp[:, ox] = np.clip(ss >> 22, 0, 255)
        for oy in range(nh):
            ymin, cnt = vb[oy]
            ss = (1 << 21) + (tmp[ymin:ymin + cnt].astype(np.int64) * np.array(vk[oy][:cnt], np.int64)[:, None]).sum(0)
            got[oy, :, c] = np.clip(ss >> 22, 0, 255)
    assert np.array_equal(got, want)


def test_fold_layernorm_algebra():
    """``kernels.fold_layernorm`` + the side-tensor layout of ``sonic_ln_side``:
    rstd * ([x | -mu_hi -mu_hi -mu_lo -mu_lo std_hi std_hi std_lo std_lo] W''^T) == LN(x) W^T + b  (fp32 emulation, CPU)."""
    from sonicdiffusionbayeslab_b200 import kernels as K

    g = torch.Generator().manual_seed(0)
    x = (2.0 + 3.0 * torch.randn(64, 320, generator=g)).bfloat16().float()
    w, b = torch.randn(96, 320, generator=g) / 18, torch.randn(96, generator=g)
    gamma, beta = 1 + 0.2 * torch.randn(320, generator=g), 0.3 * torch.randn(320, generator=g)
    w2 = K.fold_layernorm(w, b, gamma, beta)
    assert w2.dtype == torch.bfloat16 and w2.shape == (96, 320 + K.LN_SIDE_COLS) and (w2[:, 328:] == 0).all()
    mean, var = x.mean(1), x.var(1, unbiased=False)
    rstd, std = (var + 1e-5).rsqrt(), (var + 1e-5).sqrt()
    (m_hi, m_lo), (d_hi, d_lo) = K._hi_lo(-mean), K._hi_lo(std)
    side = torch.zeros(64, K.LN_SIDE_COLS)
    for col, v in enumerate((m_hi, m_hi, m_lo, m_lo, d_hi, d_hi, d_lo, d_lo)):
        side[:, col] = v.float()
    got = rstd[:, None] * (torch.cat([x, side], 1) @ w2.float().t())
    want = torch.nn.functional.layer_norm(x, (320,), gamma, beta, 1e-5) @ w.t() + b
    assert (got - want).abs().max().item() < 2e-2 * want.abs().max().item()      # only the bf16 rounding of gamma.W
    exact = torch.nn.functional.layer_norm(x, (320,), torch.ones(320), torch.zeros(320), 1e-5) @ w2[:, :320].float().t() \
        + (w.float() @ beta + b)
    assert (got - exact).abs().max().item() < 2e-4 * exact.abs().max().item()    # the identity itself (hi/lo splits)


def test_gemm_tile_width_choices_for_the_cta_pair_kernel():
    """`sonic_gemm_choose_block_n` (csrc/gemm.cu pick_block_n) without a GPU assumes 148 SMs = 74 CTA pairs: the tile
    widths of the UNet's layer shapes at UNet batch 32 follow waves x (bn / 2 + ~48 cycles per K step) -- wide tiles
    where the last wave stays full enough, narrow ones for tiny problems; GEGLU tiles hold a value and a gate half."""
    from sonicdiffusionbayeslab_b200 import kernels as k

    cases = {
        (320, 32, 64, 64, k.EPI_NONE): 160,       # the only even split of 320 below 256
        (640, 32, 32, 32, k.EPI_NONE): 160,       # 512 items on 74 clusters: 7 waves of 160 beat 6 of 256
        (1920, 1, 1, 32768, k.EPI_NONE): 256,     # QKV of the 32x32 level
        (3840, 1, 1, 8192, k.EPI_NONE): 256,      # QKV of the 16x16 level
        (2560, 1, 1, 131072, k.EPI_GEGLU): 256,   # 128 value + 128 gate columns per tile
        (16, 32, 64, 64, k.EPI_NONE): 16,         # the padded 4 -> 16 conv_out
    }
    for (N, n_img, H, W, epi), want in cases.items():
        assert k.gemm_block_n(N, n_img, H, W, epi) == want, (N, n_img, H, W, epi)
    for N in (320, 640, 960, 1280, 1920, 2560, 3840):
        bn = k.gemm_block_n(N, 32, 16, 16)
        assert bn % 32 == 0 and 32 <= bn <= 256          # staged epilogue chunks, pairable halves of whole swizzle atoms


def test_bench_cpu_sample_sizing():
    """bench.py sizes the CPU arm's timed sample from a calibration step: a whole 25-step image when it fits the
    budget (no extrapolation), else the first n steps; the line's text says which."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench._sample_size(1.01, 27.0) == 25 and bench._sample_size(1.2, 27.0) == 22
    assert bench._sample_size(3.9, 15.0) == 3 and bench._sample_size(100.0, 15.0) == 1
    assert "whole 25-step image" in bench._sample_text(25, 1.0) and "extrapolated" not in bench._sample_text(25, 1.0)
    assert "first 4 of the 25" in bench._sample_text(4, 3.9) and "extrapolated x25/4" in bench._sample_text(4, 3.9)


def test_skip_steps_example_config_resolves():
    """configs/skip_steps_config.yaml (the reference ships none for skip_steps_exp.py): names resolve through the
    registries, the zipped lists have one skip list per sweep point, and the driver builds the scheduler with the
    solver keys of skip_steps_exp.py:21-26."""
    import os

    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.config import load as load_config
    from sonicdiffusionbayeslab_b200.registry import methods_registry, models_registry, schedulers_registry

    import sonicdiffusionbayeslab_b200.experiments  # noqa: F401

    cfg = load_config(os.path.join(ROOT, "configs", "skip_steps_config.yaml"))
    assert models_registry[cfg.model.model_name] is M.StableDiffusionModelSkipTimesteps
    assert methods_registry[cfg.experiment.method].__name__ == "SkipStepsMethod"
    assert schedulers_registry[cfg.scheduler.scheduler_name] is S.DPMSolverScheduler
    p = cfg.experiment_params
    assert len(p.skip_steps) == len(p.num_inference_steps) == 2
    assert [list(v) for v in p.skip_steps] == [[3, 7, 11], [2, 4, 6, 8, 10]]
    assert all(max(s) < n for s, n in zip(p.skip_steps, p.num_inference_steps))
    assert (p.solver_order, p.algorithm_type, p.final_sigmas_type) == (2, "dpmsolver++", "zero")


def test_literal_float_rescale_constants_match_hf_processor():
    """``--literal_float_rescale`` of calc_clip_score.py (SURVEY C-8): the reference script hands float [0,1] tensors
    to torchmetrics' CLIPScore, whose HF processor rescales them by 1/255 a second time.  The product reproduces that
    with the SAME preprocessing kernel and (255 mean, 255 std); checked here against ``CLIPImageProcessor`` itself
    on natural-image-like (smooth) inputs to two grey levels / 65025 / std: HF resizes float inputs without the uint8
    rounding and clamping of the PIL path, so only bicubic overshoot on noise-like images would differ by more."""
    import warnings

    from sonicdiffusionbayeslab_b200 import kernels as K
    from transformers import CLIPImageProcessor

    proc = CLIPImageProcessor(do_resize=True, size={"shortest_edge": 224}, resample=3, do_center_crop=True,
                              crop_size={"height": 224, "width": 224}, do_rescale=True, rescale_factor=1 / 255,
                              do_normalize=True, image_mean=list(K.CLIP_MEAN), image_std=list(K.CLIP_STD),
                              do_convert_rgb=True)
    g = torch.Generator().manual_seed(0)
    low = torch.rand(2, 3, 12, 15, generator=g)
    u8 = (torch.nn.functional.interpolate(low, size=(256, 320), mode="bicubic", align_corners=False).clamp(0, 1)
          * 255).to(torch.uint8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        std_out = proc(images=[i for i in u8], return_tensors="pt")["pixel_values"]
        lit_out = proc(images=[i for i in u8.float() / 255], return_tensors="pt")["pixel_values"]
    m1, s1 = (torch.tensor(v).view(1, 3, 1, 1) for v in K.clip_norm_constants(False))
    m2, s2 = (torch.tensor(v).view(1, 3, 1, 1) for v in K.clip_norm_constants(True))
    v_over_255 = std_out * s1 + m1                       # what the kernel holds before normalising
    assert (lit_out - (v_over_255 - m2) / s2).abs().max().item() < 2.0 / (65025 * min(K.CLIP_STD))
    assert (lit_out - std_out).abs().max().item() > 1.0  # ... and it is a different image (near black)


def test_pipeline_call_argument_checks():
    """The head of the reference ``call`` (src/models.py:64-114 -> diffusers ``check_inputs``): same conditions and
    ``ValueError``s; arguments the engine does not implement raise instead of being ignored; everything is decided
    before the first CUDA access."""
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer

    pipes = [cls({}, vae=None, text_encoder=None, tokenizer=HashTokenizer(),
                 scheduler=S.DDIMSchedulerMy.from_config(M.SD15_SCHEDULER_CONFIG))
             for cls in (M.StableDiffusionModel, M.StableDiffusionModelSkipTimesteps, M.StableDiffusionModelTwoSchedulers)]
    pipes[2].scheduler_first = S.DDIMSchedulerMy.from_config(M.SD15_SCHEDULER_CONFIG)
    pipes[2].scheduler_second = S.DPMSolverScheduler.from_config(M.SD15_SCHEDULER_CONFIG)
    pe, ne = torch.zeros(2, 77, 768), torch.zeros(2, 77, 768)
    lat = torch.zeros(2, 4, 64, 64)
    for pipe in pipes:
        ok = dict(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, output_type="latent")
        for bad, exc, frag in [
            (dict(prompt_embeds=None), ValueError, "Provide either `prompt` or `prompt_embeds`"),
            (dict(prompt=["a", "b"]), ValueError, "Cannot forward both `prompt`"),
            (dict(prompt=("a", "b"), prompt_embeds=None), ValueError, "has to be of type `str` or `list`"),
            (dict(height=510, width=512), ValueError, "divisible by 8"),
            (dict(height=768, width=768), NotImplementedError, "latent_size"),
            (dict(negative_prompt="x"), ValueError, "Cannot forward both `negative_prompt`"),
            (dict(negative_prompt_embeds=ne[:1]), ValueError, "must have the same shape"),
            (dict(sigmas=[1.0, 0.5]), ValueError, "custom sigmas"),
            (dict(sigmas=[1.0], timesteps=[5]), ValueError, "Only one of `timesteps` or `sigmas`"),
            (dict(num_images_per_prompt=0), ValueError, "num_images_per_prompt"),
            (dict(num_images_per_prompt=2), ValueError, "Unexpected latents shape"),     # 2 prompts x 2 -> 4 latents
            (dict(clip_skip=1), NotImplementedError, "clip_skip"),
            (dict(callback_steps=0), ValueError, "callback_steps"),
            (dict(callback_on_step_end_tensor_inputs=["latents", "nope"]), ValueError, "tensor_inputs"),
            (dict(output_type="jpeg"), ValueError, "output_type"),
            (dict(latents=lat[:, :, :32]), ValueError, "Unexpected latents shape"),
        ]:
            with pytest.raises(exc, match=frag.replace("(", r"\\(")):
                pipe(**{**ok, **bad})
        with pytest.raises(RuntimeError, match="CUDA"):         # valid arguments: only now is the device needed
            pipe(**ok)
        with pytest.raises(RuntimeError, match="CUDA"):         # guidance_rescale is implemented (models.py:244-250)
            pipe(**ok, guidance_rescale=0.7)
        assert pipe.guidance_rescale == 0.7
        with pytest.raises(RuntimeError, match="CUDA"):         # any subset of the three names (models.py:263-267)
            pipe(**ok, callback_on_step_end_tensor_inputs=["prompt_embeds", "negative_prompt_embeds"])
        with pytest.raises(RuntimeError, match="CUDA"):         # num_images_per_prompt is implemented (models.py:173)
            pipe(**{**ok, "latents": torch.zeros(6, 4, 64, 64)}, num_images_per_prompt=3)


def test_postprocess_images_follows_diffusers():
    """``VaeImageProcessor.postprocess`` (src/models.py:313-315) on denormalised tensors: pt / np / pil."""
    import numpy as np

    from sonicdiffusionbayeslab_b200.models import postprocess_images

    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 8, 8, generator=g)
    assert postprocess_images(x, "pt") is x
    arr = postprocess_images(x, "np")
    assert arr.shape == (2, 8, 8, 3) and arr.dtype == np.float32 and np.array_equal(arr, x.permute(0, 2, 3, 1).numpy())
    pil = postprocess_images(x, "pil")
    assert len(pil) == 2 and pil[0].size == (8, 8) and pil[0].mode == "RGB"
    assert np.array_equal(np.asarray(pil[1]), (arr[1] * 255).round().astype("uint8"))
    with pytest.raises(ValueError):
        postprocess_images(x, "jpeg")


def test_stock_scheduler_is_read_from_a_local_model_directory(tmp_path):
    """``from_pretrained`` leaves the model directory's own scheduler in ``pipe.scheduler`` (diffusers reads
    ``scheduler/scheduler_config.json``; the reference's ``default`` / ``deep_cache`` methods run with it and every
    other method feeds its config to ``from_config``, base_experiment.py:69-72)."""
    import json
    import warnings

    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    none = M.load_stock_scheduler(None)
    assert type(none) is S.PNDMScheduler and dict(none.config) == dict(
        S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG).config)
    assert type(M.load_stock_scheduler(str(tmp_path))) is S.PNDMScheduler          # directory without a scheduler/

    (tmp_path / "scheduler").mkdir()
    sd15 = {"_class_name": "PNDMScheduler", "_diffusers_version": "0.6.0", "beta_end": 0.012,
            "beta_schedule": "scaled_linear", "beta_start": 0.00085, "num_train_timesteps": 1000,
            "set_alpha_to_one": False, "skip_prk_steps": True, "steps_offset": 1, "trained_betas": None,
            "clip_sample": False}
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(sd15))
    s = M.load_stock_scheduler(str(tmp_path))
    assert type(s) is S.PNDMScheduler and dict(s.config) == dict(none.config)      # the shipped SD-v1.5 file
    assert "_class_name" not in s.config and "_diffusers_version" not in s.config

    ddim = dict(sd15, _class_name="DDIMScheduler", beta_schedule="linear", beta_start=0.0001, beta_end=0.02,
                steps_offset=0)
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(ddim))
    s = M.load_stock_scheduler(str(tmp_path))
    assert type(s) is S.DDIMSchedulerMy and s.config.beta_schedule == "linear" and s.config.steps_offset == 0
    assert torch.allclose(s.betas[[0, -1]], torch.tensor([0.0001, 0.02]))

    # a stock class without a fused step (Lykon/dreamshaper-7 ships DEIS): its config survives for ``from_config``
    deis = dict(sd15, _class_name="DEISMultistepScheduler", solver_order=2, algorithm_type="deis", solver_type="logrho",
                lower_order_final=True, thresholding=False, prediction_type="epsilon", timestep_spacing="leading")
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(deis))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        s = M.load_stock_scheduler(str(tmp_path))
    assert any("DEISMultistepScheduler" in str(x.message) for x in w)
    assert type(s) is S.PNDMScheduler and s.config.solver_order == 2 and s.config.algorithm_type == "deis"
    lcm = S.LCMScheduler.from_config(s.config)                                     # consistency_model.py's idiom
    assert lcm.config.beta_start == 0.00085 and lcm.config.steps_offset == 1
    # a known class with an option that has no fused step: same stand-in, config kept, loud
    karras = dict(sd15, _class_name="DPMSolverMultistepScheduler", use_karras_sigmas=True, solver_order=2)
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(karras))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        s = M.load_stock_scheduler(str(tmp_path))
    assert any("not fused here" in str(x.message) for x in w)
    assert type(s) is S.PNDMScheduler and s.config.use_karras_sigmas is True and s.config.solver_order == 2


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The four argument structs of include/sonic.h as the C compiler lays them out (gcc, the platform ABI nvcc's host
    side uses too) against the ``ctypes.Structure`` mirrors the Python host binds with (``_lib.GemmArgs``,
    ``kernels.AttentionArgs / UpdateCoeffs / X0Post``): same size, same field names in the same order, same offsets
    -- ctypes checks none of this at call time."""
    import ctypes
    import json
    import re
    import subprocess

    from sonicdiffusionbayeslab_b200 import _lib
    from sonicdiffusionbayeslab_b200 import kernels as K

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sonic.h")).read()
    mirrors = {"sonic_gemm_args": _lib.GemmArgs, "sonic_attention_args": K.AttentionArgs,
               "sonic_update_coeffs": K.UpdateCoeffs, "sonic_x0_post": K.X0Post}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sonic.h"', "int main(void) {", '  printf("{");']
    for si, (name, mirror) in enumerate(mirrors.items()):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):                       # "const void* a0", "int32_t c0, ld0", "float m_x, m_e" ...
            decl = decl.strip()
            if decl:
                first, *rest = decl.split(",")
                fields += [first.split()[-1].lstrip("*")] + [r.strip().lstrip("*") for r in rest]
        assert fields == [f[0] for f in mirror._fields_], (name, fields)          # names and order
        lines.append('  printf("%s\\"%s\\": {\\"sizeof\\": %%zu", sizeof(%s));' % (", " if si else "", name, name))
        for f in fields:
            lines.append('  printf(", \\"%s\\": %%zu", offsetof(%s, %s));' % (f, name, f))
        lines.append('  printf("}");')
    lines += ['  printf("}\\n");', "  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    layout = json.loads(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    for name, mirror in mirrors.items():
        assert ctypes.sizeof(mirror) == layout[name]["sizeof"], name
        for f, _ in mirror._fields_:
            assert getattr(mirror, f).offset == layout[name][f], (name, f)


def test_every_exported_entry_point_is_declared_in_the_header():
    """The reverse of ``test_library_exports_every_declared_symbol``: libsonic.so exports no ``sonic_*`` entry point
    that include/sonic.h does not declare (the clock-trace hooks ``sonic_debug_*`` of tools/ excepted)."""
    from sonicdiffusionbayeslab_b200 import _lib

    try:
        out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True,
                             check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("nm unavailable")
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("sonic_")}
    undeclared = sorted(n for n in exported - set(_declared_symbols()) if not n.startswith("sonic_debug_"))
    assert not undeclared, undeclared


def test_call_sites_pass_as_many_arguments_as_the_header_declares():
    """ctypes checks neither the number nor the width of the arguments: every ``lib.sonic_*(...)`` call in the package,
    bench.py and calc_clip_score.py is counted against the prototype in include/sonic.h, and every ``int64_t``
    parameter must be passed as an explicit ``c_int64`` (a bare Python int goes out as a 32-bit C int)."""
    import ast
    import pathlib
    import re

    root = pathlib.Path(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    hdr = re.sub(r"/\*.*?\*/", "", (root / "include" / "sonic.h").read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(sonic_\w+)\s*\(([^;{]*?)\)\s*;", hdr, re.S):
        params = m.group(2).strip()
        protos[m.group(1)] = [] if params in ("void", "") else [" ".join(p.split()) for p in params.split(",")]
    assert set(protos) == set(_declared_symbols())
    files = list((root / "sonicdiffusionbayeslab_b200").rglob("*.py")) + [root / "bench.py", root / "calc_clip_score.py",
                                                                         root / "__graft_entry__.py"]
    seen, problems = set(), []
    for f in files:
        for n in ast.walk(ast.parse(f.read_text())):
            if not (isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr.startswith("sonic_")):
                continue
            name, where = n.func.attr, f"{f.relative_to(root)}:{n.lineno}"
            seen.add(name)
            if name not in protos:
                problems.append(f"{where}: {name} is not declared in include/sonic.h")
                continue
            if n.keywords or any(isinstance(a, ast.Starred) for a in n.args):
                problems.append(f"{where}: {name} must be called with plain positional arguments")
                continue
            if len(n.args) != len(protos[name]):
                problems.append(f"{where}: {name} gets {len(n.args)} arguments, the header declares {len(protos[name])}")
                continue
            for arg, decl in zip(n.args, protos[name]):
                if decl.startswith("int64_t") and "c_int64" not in ast.unparse(arg):
                    problems.append(f"{where}: {name}: `{decl}` is passed `{ast.unparse(arg)}` (not a c_int64)")
    assert not problems, "\n".join(problems)
    assert len(seen) >= 30                                   # the check really saw the binding's call sites


def test_lora_merge_peft_and_kohya_layouts(tmp_path):
    """``load_lora_weights`` + ``fuse_lora`` (consistency_model.py:20-21): W' = W + (alpha / r) * scale * B A for the
    diffusers / PEFT key layout and for the kohya layout the LCM-LoRA ships in (conv adapters included, text-encoder
    adapters skipped); an adapter that matches nothing raises, a hub id (no network) warns and changes nothing."""
    import warnings

    from safetensors.torch import save_file

    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer

    g = torch.Generator().manual_seed(0)
    lin = "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q"
    conv = "down_blocks.0.resnets.0.conv1"
    base = {lin + ".weight": torch.randn(32, 24, generator=g), conv + ".weight": torch.randn(16, 8, 3, 3, generator=g),
            conv + ".bias": torch.randn(16, generator=g)}

    def pipe():
        return M.StableDiffusionModel({k: v.clone() for k, v in base.items()}, vae=None, text_encoder=None,
                                      tokenizer=HashTokenizer(), scheduler=S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG))

    r = 4
    a_lin, b_lin = torch.randn(r, 24, generator=g), torch.randn(32, r, generator=g)
    a_conv, b_conv = torch.randn(r, 8, 3, 3, generator=g), torch.randn(16, r, 1, 1, generator=g)
    want_lin = lambda s: base[lin + ".weight"] + s * (b_lin @ a_lin)                                   # noqa: E731
    want_conv = lambda s: base[conv + ".weight"] + s * (b_conv.flatten(1) @ a_conv.flatten(1)).view(16, 8, 3, 3)  # noqa: E731

    peft = tmp_path / "peft.safetensors"
    save_file({f"unet.{lin}.lora_A.weight": a_lin, f"unet.{lin}.lora_B.weight": b_lin,
               f"unet.{conv}.lora_A.weight": a_conv, f"unet.{conv}.lora_B.weight": b_conv}, str(peft))
    p = pipe()
    p.load_lora_weights(str(peft))
    p.fuse_lora()
    assert torch.allclose(p._unet_sd[lin + ".weight"], want_lin(1.0), atol=1e-5)
    assert torch.allclose(p._unet_sd[conv + ".weight"], want_conv(1.0), atol=1e-5)
    assert torch.equal(p._unet_sd[conv + ".bias"], base[conv + ".bias"]) and p._engines == {} and p._weights is None

    kohya = tmp_path / "kohya.safetensors"
    alpha = 8.0                                                       # alpha / r = 2
    save_file({f"lora_unet_{lin.replace('.', '_')}.lora_down.weight": a_lin,
               f"lora_unet_{lin.replace('.', '_')}.lora_up.weight": b_lin,
               f"lora_unet_{lin.replace('.', '_')}.alpha": torch.tensor(alpha),
               f"lora_unet_{conv.replace('.', '_')}.lora_down.weight": a_conv,
               f"lora_unet_{conv.replace('.', '_')}.lora_up.weight": b_conv,
               f"lora_unet_{conv.replace('.', '_')}.alpha": torch.tensor(alpha),
               "lora_te_text_model_encoder_layers_0_mlp_fc1.lora_down.weight": torch.randn(r, 8, generator=g),
               "lora_te_text_model_encoder_layers_0_mlp_fc1.lora_up.weight": torch.randn(8, r, generator=g)}, str(kohya))
    p = pipe()
    p.load_lora_weights(str(kohya), adapter_scale=0.5)
    p.fuse_lora(lora_scale=0.5)
    s = (alpha / r) * 0.5 * 0.5
    assert torch.allclose(p._unet_sd[lin + ".weight"], want_lin(s), atol=1e-5)
    assert torch.allclose(p._unet_sd[conv + ".weight"], want_conv(s), atol=1e-5)

    other = tmp_path / "other.safetensors"
    save_file({"unet.mid_block.nothing_here.lora_A.weight": a_lin, "unet.mid_block.nothing_here.lora_B.weight": b_lin},
              str(other))
    p = pipe()
    p.load_lora_weights(str(other))
    with pytest.raises(ValueError, match="none of the 2 tensors"):
        p.fuse_lora()

    p = pipe()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p.load_lora_weights("latent-consistency/lcm-lora-sdv1-5")
    assert any("NO adapter is applied" in str(x.message) for x in w)
    p.fuse_lora()                                                     # nothing loaded: a no-op, weights untouched
    assert all(torch.equal(p._unet_sd[k], v) for k, v in base.items())
