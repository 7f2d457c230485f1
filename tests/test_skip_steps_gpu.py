"""The skip-timesteps pipeline on the B200 engine against the fp32 oracle (SURVEY 8(f)-4;
/root/reference/src/models.py:1220-1223,1338-1340, driver skip_steps_exp.py:55-62): the registered
``stable_diffusion_model_skip_timesteps`` class, DPM-Solver++(2M) on a 20-step grid with loop indices
2, 3, 9, 15, 16 skipped, teacher-forced on the unit-variance fixture of tests/parity_lib.py.

Gates: (i) bit-exact structure -- the UNet sees exactly the grid timesteps of the executed indices, the callback gets
LOOP indices, ``num_timesteps`` is the grid length, one UNet evaluation per executed step; (ii) every executed
step's latents within 1.3 x stock-PyTorch-bf16's worst step + 3e-2 of the fp32 oracle (the same comparison
tests/test_parity_abs_gpu.py makes for the unskipped loop: a bf16 UNet under CFG 7.5 is what loses ~1e-1 per step).
A wrong solver interval or a stale history entry -- what a skipping bug would produce -- shows up at this fixture's
magnitudes (|x| ~ 5) well above that.  (A CPU rehearsal of this very test over interpreted plans at a 16 x 16 latent --
tools/plan_interp.py, bf16 activations -- gave engine-like worst 1.24e-1, stock-PyTorch-bf16 worst 1.13e-1; at the
real 64 x 64 latent 1.06e-1 against 1.12e-1: profiles/r2m_cpu_rehearsal_skip_steps.txt.)  The host
logic of the same loop is pinned to 5e-6 against the reference's own source on the CPU
(tests/test_pipeline_host_cpu.py, case ``skip_dpmpp``).
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_lib as PL  # noqa: E402
import skip_case  # noqa: E402

pytestmark = pytest.mark.gpu

N_STEPS, SKIP = 20, [2, 3, 9, 15, 16]


def test_skip_timesteps_pipeline_vs_fp32_oracle(cuda):
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200.registry import models_registry

    cls = models_registry["stable_diffusion_model_skip_timesteps"]
    assert cls is M.StableDiffusionModelSkipTimesteps
    net, net16, _ = PL.unit_variance_unet(cuda)
    model = PL.make_model(dict(net.state_dict()), cuda, cls)
    pe, ne, z0, noise = PL.inputs(cuda, 2)
    r = skip_case.run(model, net, net16, pe, ne, z0, noise, N_STEPS, SKIP)
    e, f = r["engine"], r["torch_bf16"]
    print(f"\n[skip-timesteps DPM++ {N_STEPS} steps, skipped {SKIP}] |x|max {r['xmax']:.2f}: engine worst "
          f"{max(e):.3e} median {sorted(e)[len(e) // 2]:.3e}; torch-bf16 worst {max(f):.3e}")
    executed = [i for i in range(N_STEPS) if i not in SKIP]
    assert [i for i, _ in r["seen"]] == executed                               # callback gets loop indices
    assert [t for _, t in r["seen"]] == [r["timesteps"][i] for i in executed]  # UNet timesteps: grid values
    assert model.num_timesteps == N_STEPS and len(model.last_step_kinds) == len(executed)
    assert r["x0"] == [] and r["secs"] > 0
    assert r["xmax"] < 10.0                                                    # SD-like magnitudes
    assert max(e) <= 1.3 * max(f) + 3e-2, (e, f)
    assert torch.isfinite(r["out"].images.float()).all()
