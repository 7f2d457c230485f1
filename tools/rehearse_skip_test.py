"""CPU rehearsal of tests/test_skip_steps_gpu.py: the same harness (tests/skip_case.py), the same unit-variance fixture
(tests/parity_lib.py), the product pipeline and a REAL ``UNetEngine`` -- recorded over host buffers and replayed by the
plan interpreter (tools/plan_interp.py: bf16 activations, fp32 accumulation) with the fused scheduler step through its
float64 model -- in place of the GPU.  Prints the per-step max-abs of the engine-like path and of stock-PyTorch bf16
against the fp32 oracle, i.e. what the GPU test gates, so its tolerance is not chosen blind.

    python tools/rehearse_skip_test.py [latent size, default 16]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

import parity_lib as PL  # noqa: E402
import plan_check  # noqa: E402
import plan_interp  # noqa: E402
import skip_case  # noqa: E402
from test_pipeline_host_cpu import _launch_in_place  # noqa: E402

from sonicdiffusionbayeslab_b200 import kernels as K  # noqa: E402
from sonicdiffusionbayeslab_b200 import models as M  # noqa: E402
from sonicdiffusionbayeslab_b200 import schedulers as S  # noqa: E402
from sonicdiffusionbayeslab_b200 import unet_engine as UE  # noqa: E402
from sonicdiffusionbayeslab_b200.text import HashTokenizer  # noqa: E402


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    net, net16, _ = PL.unit_variance_unet("cpu")
    packed = UE.PackedWeights(dict(net.state_dict()), "cpu")
    pe, ne, z0, noise = PL.inputs("cpu", 2)
    z0, noise = z0[:, :, :hw, :hw].contiguous(), noise[:, :, :hw, :hw].contiguous()
    interp = plan_interp.PlanInterpreter()
    with plan_check.recording() as tr:
        record = tr.on_op

        def on_op(name, args):
            interp.record(name, args)
            record(name, args)

        tr.on_op = on_op
        UE._Plan.run = lambda plan, stream: interp.run(plan.h)
        K.stream_ptr = lambda: None
        S.FusedScheduler._launch = _launch_in_place
        torch.cuda.synchronize = lambda *a, **k: None

        def engine(self, n_latents, cfg_dup):
            key = (n_latents, cfg_dup)
            if key not in self._engines:
                tr.raws.clear()
                self._engines[key] = UE.UNetEngine(packed, n_latents=n_latents, cfg_dup=cfg_dup, height=hw, width=hw,
                                                   io_dtype=self.dtype, device="cpu")
            return self._engines[key]

        M._PipelineBase.engine = engine
        model = M.StableDiffusionModelSkipTimesteps(
            packed.sd, vae=None, text_encoder=None, tokenizer=HashTokenizer(),
            scheduler=S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG), torch_dtype=torch.bfloat16, latent_size=hw)
        t0 = time.time()
        r = skip_case.run(model, net, net16, pe, ne, z0, noise, 20, [2, 3, 9, 15, 16])
        e, f = r["engine"], r["torch_bf16"]
        print(f"latent {hw}x{hw}, {time.time() - t0:.0f} s: engine-like worst {max(e):.3e} median "
              f"{sorted(e)[len(e) // 2]:.3e} | stock-PyTorch bf16 worst {max(f):.3e} | |x|max {r['xmax']:.2f} | "
              f"plan problems {len(tr.problems)}")
        print("engine-like per step:", " ".join(f"{v:.2e}" for v in e))
        print("stock bf16 per step :", " ".join(f"{v:.2e}" for v in f))
        print(f"gate of the GPU test: 1.3 x {max(f):.3e} + 3e-2 = {1.3 * max(f) + 3e-2:.3e}")


if __name__ == "__main__":
    main()
