"""ABSOLUTE-tolerance parity on the unit-variance fixture (BASELINE.json north_star: "latents must match per step
within a stated tolerance, e.g. max-abs <= 2e-2 in bf16"), plus the direct UNet test and the batch-32 oracle run.

Stated tolerances (measured numbers: profiles/r2_parity_abs.txt, written by tools/parity_report.py):
  * UNet eps, engine (bf16 tensor cores) vs fp32 oracle, eps ~ N(0,1):  max-abs <= EPS_TOL, and no worse than
    1.3x what stock PyTorch bf16 (cuDNN / cuBLAS / SDPA) loses on the same weights;
  * one denoising step (UNet + CFG 7.5 + scheduler), teacher-forced from the oracle, latents |x| <~ 5:
        fp32 latent I/O : max-abs <= STEP_TOL_F32
        bf16 latent I/O : max-abs <= STEP_TOL_BF16 -- one bf16 rounding of |x| in [4, 8) alone is 1.6e-2, and
                          classifier-free guidance 7.5 multiplies the UNet's bf16 eps error by ~10 before the
                          scheduler scales it, so the north_star's example figure (2e-2) is met by the median
                          step but not by the worst one; the gate is the measured worst case + 25 %, and the
                          engine must not be worse than 1.3x stock PyTorch bf16 teacher-forced the same way.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_lib as PL  # noqa: E402

pytestmark = pytest.mark.gpu

EPS_TOL = 6e-2
STEP_TOL_F32 = 5e-2
STEP_TOL_BF16 = 8e-2
MEDIAN_TOL_BF16 = 2e-2


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
    assert err <= EPS_TOL and err <= 1.3 * floor + 5e-3, (err, floor)


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
    if io == torch.float32:
        assert max(e) <= STEP_TOL_F32, e
    else:
        assert max(e) <= STEP_TOL_BF16 and med <= MEDIAN_TOL_BF16, e
        if f:
            assert max(e) <= 1.3 * max(f) + 5e-3, (max(e), max(f))


def test_batch32_three_dpm_steps_vs_fp32_oracle(unit):
    """The BENCHMARKED shape (batch 16, CFG -> UNet batch 32) against the fp32 oracle run on the GPU: the first
    three steps of the genuine 25-step DPM-Solver++(2M) schedule (order 1, then 2, 2), teacher-forced."""
    r = PL.teacher_forced("dpmpp25", unit["net"], None, unit["sd"], unit["dev"], io_dtype=torch.bfloat16, B=16,
                          max_steps=3)
    print(f"\n[batch 16 / UNet batch 32, DPM++ steps 1-3 vs fp32 oracle] max-abs {r['engine']}, |x|max {max(r['xmax']):.2f}")
    assert len(r["engine"]) == 3 and max(r["xmax"]) < 8.0
    assert max(r["engine"]) <= STEP_TOL_BF16, r["engine"]
