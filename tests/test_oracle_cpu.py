"""CPU tests: the oracle against the golden known answers (tests/golden/schedule_kat.json) and its
own structural invariants.  Integer schedules are bit-exact; float32 tables agree to 2e-5 relative
with the independent numpy derivation of tests/golden/make_golden.py."""
import json
import os

import numpy as np
import pytest
import torch

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "schedule_kat.json")))


def _sd15(cls, **kw):
    from oracle.schedulers import sd15

    return sd15(cls, **kw)


def test_alphas_cumprod_and_ddim_grid():
    from oracle.schedulers import DDIMScheduler

    s = _sd15(DDIMScheduler)
    for i, v in KAT["alphas_cumprod"].items():
        assert abs(float(s.alphas_cumprod[int(i)]) / v - 1) < 2e-5
    s.set_timesteps(20)
    assert s.timesteps.tolist() == KAT["ddim_20"]
    s.set_timesteps(50)
    assert s.timesteps.tolist() == KAT["ddim_50"]
    assert abs(float(s.alphas_cumprod[951]) / KAT["ddim_951"]["a_t"] - 1) < 2e-5
    assert float(s.final_alpha_cumprod) == float(s.alphas_cumprod[0])     # set_alpha_to_one=False


def test_dpm_grid_and_sigmas():
    from oracle.schedulers import DPMSolverScheduler

    s = _sd15(DPMSolverScheduler, solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")
    s.set_timesteps(25)
    assert s.timesteps.tolist() == KAT["dpm_25"]
    assert np.allclose(s.sigmas[:3].tolist(), KAT["dpm_25_sigmas_head"], rtol=2e-5)
    assert np.allclose(s.sigmas[-3:].tolist(), KAT["dpm_25_sigmas_tail"], rtol=2e-5, atol=1e-9)
    assert s.sigmas.dtype == torch.float32 and s.sigmas[-1] == 0
    m = _sd15(DPMSolverScheduler, algorithm_type="dpmsolver", final_sigmas_type="sigma_min")
    m.set_timesteps(25)
    assert abs(float(m.sigmas[-1]) / KAT["sigma_min"] - 1) < 2e-5
    with pytest.raises(ValueError):
        _sd15(DPMSolverScheduler, algorithm_type="dpmsolver", final_sigmas_type="zero")
    with pytest.raises(NotImplementedError):
        _sd15(DPMSolverScheduler, algorithm_type="")


def test_lcm_and_pndm_grids():
    from oracle.schedulers import LCMScheduler, PNDMScheduler

    s = _sd15(LCMScheduler)
    for n in (1, 2, 4):
        s.set_timesteps(n)
        assert s.timesteps.tolist() == KAT[f"lcm_{n}"]
    with pytest.raises(ValueError):
        s.set_timesteps(51)
    c_skip, c_out = s.get_scalings_for_boundary_condition_discrete(torch.tensor(999))
    assert abs(float(c_skip) / KAT["lcm_c_skip_999"] - 1) < 1e-5 and abs(float(c_out) - 1) < 1e-6
    p = _sd15(PNDMScheduler)
    p.set_timesteps(50)
    assert p.timesteps.tolist() == KAT["pndm_50"]


def test_dpm_final_step_returns_x0_and_order_flags():
    from oracle.schedulers import DPMSolverScheduler

    s = _sd15(DPMSolverScheduler, solver_order=2, algorithm_type="dpmsolver++", final_sigmas_type="zero")
    s.set_timesteps(25)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 8, 8, generator=g)
    orders = []
    for t in s.timesteps:
        before = s.lower_order_nums
        last = s.step_index is not None and s.step_index == len(s.timesteps) - 1
        prev, x0 = s.step(torch.randn(1, 4, 8, 8, generator=g), t, x)
        orders.append(1 if before < 1 or last else 2)
        x = prev
    assert orders[0] == 1 and orders[1] == 2
    assert torch.allclose(prev, x0, atol=1e-5)          # sigma_last = 0 -> x' = x0 (appendix A.2.2)
    assert s.step_index == 25


def test_two_scheduler_switch_lists():
    from oracle.pipeline import switch_timestamp
    from oracle.schedulers import DDIMScheduler, DPMSolverScheduler

    for (n1, k), key in (((10, 3), "two_10_3"), ((20, 10), "two_20_10")):
        a, b = _sd15(DDIMScheduler), _sd15(DPMSolverScheduler)
        a.set_timesteps(n1)
        b.set_timesteps(timesteps=a.timesteps.numpy())
        first, second = switch_timestamp(a.timesteps, b.timesteps, k, "closest")
        assert [int(t) for t in first] == KAT[key][0] and [int(t) for t in second] == KAT[key][1]
        for mode in ("left_closest", "right_closest"):    # shared grid: all modes agree (appendix C-6)
            f2, s2 = switch_timestamp(a.timesteps, b.timesteps, k, mode)
            assert [int(t) for t in s2] == KAT[key][1]


def test_unet_param_count_and_temb():
    from oracle.unet import UNet2DConditionModel, timestep_embedding

    with torch.device("meta"):
        net = UNet2DConditionModel()
    assert sum(p.numel() for p in net.parameters()) == KAT["unet_params"]
    e = timestep_embedding(torch.tensor([951]), 320)[0, [0, 1, 160, 161]].tolist()
    assert np.allclose(e, KAT["temb_951"], atol=2e-4)


def _small():
    from oracle.unet import UNetConfig, make_unet

    return make_unet(cfg=UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32, num_heads=2))


def test_small_unet_pipeline_and_deepcache_semantics():
    from oracle.deepcache import DeepCacheOracle
    from oracle.pipeline import denoise
    from oracle.schedulers import DDIMScheduler, PNDMScheduler

    net = _small()
    g = torch.Generator().manual_seed(1)
    pe, ne = torch.randn(2, 7, 32, generator=g), torch.randn(2, 7, 32, generator=g)
    x = torch.randn(2, 4, 16, 16, generator=g)
    plain = denoise(net, _sd15(DDIMScheduler), pe, ne, x, 4)
    dc = DeepCacheOracle(net)
    dc.set_params(cache_interval=1, cache_branch_id=0)
    same = denoise(net, _sd15(DDIMScheduler), pe, ne, x, 4, deepcache=dc)
    assert torch.equal(plain["latents"], same["latents"])            # interval 1 == no caching
    assert len(plain["x0"]) == 4 and plain["x0"][0].shape == (1, 4, 16, 16)
    dc.set_params(cache_interval=3, cache_branch_id=0)
    cached = denoise(net, _sd15(PNDMScheduler), pe, ne, x, 5, deepcache=dc)
    assert len(cached["timesteps"]) == 6 and len(cached["x0"]) == 0  # PLMS: N+1 UNet calls, 1-tuple steps
    assert not torch.equal(cached["latents"], denoise(net, _sd15(PNDMScheduler), pe, ne, x, 5)["latents"])
    # teacher forcing reproduces the free run when fed its own trajectory
    forced = [x] + plain["per_step"][:-1]
    again = denoise(net, _sd15(DDIMScheduler), pe, ne, x, 4, forced_latents=forced)
    assert torch.equal(again["latents"], plain["latents"])


def test_lcm_consumes_generator_in_order():
    from oracle.pipeline import denoise
    from oracle.schedulers import LCMScheduler

    net = _small()
    g = torch.Generator().manual_seed(2)
    pe = torch.randn(1, 7, 32, generator=g)
    x = torch.randn(1, 4, 16, 16, generator=g)
    a = denoise(net, _sd15(LCMScheduler), pe, pe, x, 4, guidance_scale=0, generator=torch.Generator().manual_seed(7))
    b = denoise(net, _sd15(LCMScheduler), pe, pe, x, 4, guidance_scale=0, generator=torch.Generator().manual_seed(7))
    c = denoise(net, _sd15(LCMScheduler), pe, pe, x, 4, guidance_scale=0, generator=torch.Generator().manual_seed(8))
    assert torch.equal(a["latents"], b["latents"]) and not torch.equal(a["latents"], c["latents"])
    assert len(a["x0"]) == 4


def test_interleaved_partition_and_history_feed():
    """Interleaved-scheduler host logic (models.py:944-961, 1010-1053) on a toy epsilon model: which timesteps are
    evaluated, which go to the inter scheduler, and that the main scheduler's history is fed after an inter step."""
    from oracle import schedulers as O
    from oracle.pipeline import denoise_interleaved, interleave_partition

    assert interleave_partition(range(12), 3, [0, 2]) == ([0, 3, 4, 5, 6, 9, 10, 11], [0, 6])
    assert interleave_partition([951, 901, 851, 801], 2, []) == ([951, 901, 851, 801], [])

    class Toy(torch.nn.Module):
        def forward(self, x, t, encoder_hidden_states=None):
            return (0.1 * x + 0.001 * float(t),)

    cfg = O.SD15_SCHEDULER_CONFIG
    lat = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    pe, ne = torch.zeros(2, 77, 8), torch.zeros(2, 77, 8)
    main, inter = O.DPMSolverScheduler.from_config(cfg), O.DDIMScheduler.from_config(cfg)
    out = denoise_interleaved(Toy(), main, inter, pe, ne, lat, 10, [1, 3], guidance_scale=7.5)
    assert out["timesteps"] == ([901, 811, 721, 541, 451, 361, 181, 91], [721, 361])
    assert len(out["per_step"]) == 8 and torch.isfinite(out["latents"]).all()
    assert main.step_index == 6                       # six main steps: its counter ignores the two inter steps
    # no interleaving == the plain DPM loop
    from oracle.pipeline import denoise
    a = denoise_interleaved(Toy(), O.DPMSolverScheduler.from_config(cfg), O.DDIMScheduler.from_config(cfg), pe, ne,
                            lat, 10, [], guidance_scale=7.5)
    b = denoise(Toy(), O.DPMSolverScheduler.from_config(cfg), pe, ne, lat, 10, guidance_scale=7.5)
    assert torch.equal(a["latents"], b["latents"])
