"""ABSOLUTE-tolerance parity on the SD-like fixture (BASELINE.json north_star: "latents must match per step within
a stated tolerance, e.g. max-abs <= 2e-2 in bf16"), plus the direct UNet test and the batch-32 oracle run.

The fixture (tests/parity_lib.py): unit-variance eps, correlated cond / uncond contexts, every step entered with a
forward-process sample (|x|max ~ 5).  MEASURED on B200 (profiles/r2_parity_abs.txt, tools/parity_report.py), max-abs
against the fp32 oracle run through stock PyTorch on the same GPU:

                               engine (bf16 I/O)     engine (fp32 I/O)     stock PyTorch bf16 (cuDNN/cuBLAS/SDPA)
    UNet eps, t = 981/501/21   3.5e-2 / 3.3e-2 / 3.2e-2     --             5.3e-2 / 4.5e-2 / 5.9e-2
    DDIM-20 step               worst 8.8e-2  median 4.5e-2   7.5e-2 / 3.8e-2      1.0e-1 / 6.3e-2
    DPM-Solver++-25 step       9.5e-2 / 4.7e-2               9.4e-2 / 4.6e-2      9.9e-2 / 5.4e-2
    PNDM-20 step               2.3e-1 / 1.2e-1               2.4e-1 / 1.2e-1      2.7e-1 / 1.5e-1
    DDIM-12 + DeepCache(3)     1.4e-1 / 6.4e-2               1.3e-1 / 6.0e-2      1.8e-1 / 8.6e-2
    two-scheduler 20 / k=10    2.0e-1 / 5.8e-2               1.9e-1 / 4.8e-2      --
    DPM++ steps 1-3, batch 16  6.2e-2, 8.9e-2, 9.5e-2

So the north_star's example figure (2e-2) is NOT met -- by the engine or by stock PyTorch bf16, with bf16 or fp32
latents: it is the bf16 UNet evaluation itself that loses ~3-5e-2 on a unit-variance eps (8-bit mantissas through
~200 layers), classifier-free guidance 7.5 turns that into ~10x on the guided eps (-6.5 u + 7.5 c) and the scheduler
scales it by 0.1-0.3 per step (PLMS by up to 55/24).  What IS met and gated here: (i) the engine's error is below
stock PyTorch bf16's in every row; (ii) absolute gates at the measured worst case + 30 %; (iii) the scheduler
arithmetic alone (fused kernel, same eps in) meets 1e-4 fp32 / 2e-2 bf16 (tests/test_pipeline_gpu.py,
tests/test_reference_pins_gpu.py) and integer schedules are bit-exact.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_lib as PL  # noqa: E402

pytestmark = pytest.mark.gpu

EPS_TOL = 4.5e-2                     # UNet eps max-abs, engine vs fp32 oracle (measured 3.2-3.5e-2)
STEP_TOL = {                         # per-step latent max-abs, measured worst case x 1.3 (bf16 and fp32 latent I/O)
    "ddim20": 1.15e-1, "dpmpp25": 1.25e-1, "pndm20": 3.1e-1, "deepcache_ddim12_i3": 1.9e-1, "two_20_k10": 2.6e-1}
VS_LIBRARY = 1.05                    # engine MEDIAN step error (and eps error) <= 1.05 x stock-PyTorch-bf16's (+ 5e-3)
VS_LIBRARY_WORST = 1.2               # the worst step of 12-25 is a noisier statistic: <= 1.2 x stock PyTorch bf16's worst


@pytest.fixture(scope="module")
def unit(cuda):
    net, net16, scale = PL.unit_variance_unet(cuda)
    return dict(net=net, net16=net16, sd=dict(net.state_dict()), dev=cuda, models={})


def _model(unit, cls, io):
    key = (cls, io)
    if key not in unit["models"]:
        unit["models"][key] = PL.make_model(unit["sd"], unit["dev"], cls, io_dtype=io)
    return unit["models"][key]


@pytest.mark.parametrize("t", [981.0, 501.0, 21.0])
def test_unet_engine_eps_matches_oracle(unit, t):
    """Direct ``UNetEngine.forward`` vs ``oracle.unet`` (no scheduler, no guidance)."""
    from sonicdiffusionbayeslab_b200.unet_engine import UNetEngine

    dev = unit["dev"]
    if "eng" not in unit:
        unit["eng"] = UNetEngine(unit["sd"], n_latents=2, cfg_dup=True, io_dtype=torch.float32, device=dev)
    eng = unit["eng"]
    pe, ne, lat, _ = PL.inputs(dev)
    ctx = torch.cat([ne, pe])
    eng.x_in.copy_(lat)
    eng.set_context(ctx.bfloat16())
    got = eng.forward(t).float().clone()
    x = torch.cat([lat, lat])
    with torch.no_grad():
        want = unit["net"](x, torch.tensor(t, device=dev), encoder_hidden_states=ctx)[0]
        lib16 = unit["net16"](x.bfloat16(), torch.tensor(t, device=dev), encoder_hidden_states=ctx.bfloat16())[0].float()
    err, floor = (got - want).abs().max().item(), (lib16 - want).abs().max().item()
    print(f"\n[unet t={t:.0f}] eps std {want.std().item():.3f} |eps|max {want.abs().max().item():.2f}: "
          f"engine max-abs {err:.3e}, torch-bf16 {floor:.3e}")
    assert 0.7 < want.std().item() < 1.4                      # the fixture really is unit-variance
    assert err <= EPS_TOL and err <= VS_LIBRARY * floor + 5e-3, (err, floor)


@pytest.mark.parametrize("name", list(PL.CASES))
@pytest.mark.parametrize("io", [torch.float32, torch.bfloat16])
def test_step_absolute_error_teacher_forced(unit, name, io):
    from sonicdiffusionbayeslab_b200 import models as M

    cls = M.StableDiffusionModelTwoSchedulers if PL.CASES[name][0] == "two" else M.StableDiffusionModel
    r = PL.teacher_forced(name, unit["net"], unit["net16"], unit["sd"], unit["dev"], io_dtype=io,
                          model=_model(unit, cls, io))
    e, f = r["engine"], r["torch_bf16"]
    med = sorted(e)[len(e) // 2]
    print(f"\n[{name} io={io}] |x|max {max(r['xmax']):.2f}: engine worst {max(e):.3e} median {med:.3e}"
          + (f"; torch-bf16 worst {max(f):.3e}" if f else ""))
    assert max(r["xmax"]) < 8.0                               # SD-like magnitudes: the absolute figure means something
    assert max(e) <= STEP_TOL[name], e
    if f:
        assert max(e) <= VS_LIBRARY_WORST * max(f) + 5e-3 and med <= VS_LIBRARY * sorted(f)[len(f) // 2] + 5e-3, (e, f)


def test_batch32_three_dpm_steps_vs_fp32_oracle(unit):
    """The BENCHMARKED shape (batch 16, CFG -> UNet batch 32) against the fp32 oracle run on the GPU: the first
    three steps of the genuine 25-step DPM-Solver++(2M) schedule (order 1, then 2, 2), teacher-forced."""
    r = PL.teacher_forced("dpmpp25", unit["net"], None, unit["sd"], unit["dev"], io_dtype=torch.bfloat16, B=16,
                          max_steps=3)
    print(f"\n[batch 16 / UNet batch 32, DPM++ steps 1-3 vs fp32 oracle] max-abs {r['engine']}, |x|max {max(r['xmax']):.2f}")
    assert len(r["engine"]) == 3 and max(r["xmax"]) < 8.0
    assert max(r["engine"]) <= STEP_TOL["dpmpp25"], r["engine"]
