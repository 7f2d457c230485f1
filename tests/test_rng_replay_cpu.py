"""RNG parity of sharded runs, host side, no GPU.

The reference consumes ONE generator over every batch and every noisy scheduler step
(/root/reference/src/experiments/base_experiment.py:51-53,149; schedulers.py:134-147).  A rank that does not
own a batch must leave the generator exactly where the single process would: ``model(..., rng_only=True)``.
The oracle loop (pinned bit-for-bit against the reference's own loop source, tests/test_reference_pins_cpu.py
case ``loop_lcm4_rng``) is the witness of what a real call consumes.
"""
import pytest
import torch


def _zero_unet(x, t, encoder_hidden_states=None, **_):
    return (torch.zeros_like(x),)


def _pipe(cls=None, dtype=torch.float32):
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S
    from sonicdiffusionbayeslab_b200.text import HashTokenizer
    from sonicdiffusionbayeslab_b200.unet_spec import unet_param_shapes

    cls = cls or M.StableDiffusionModel
    sd = {k: None for k in unet_param_shapes()}         # never packed: rng_only touches no weights
    return cls(sd, vae=None, text_encoder=None, tokenizer=HashTokenizer(),
               scheduler=S.PNDMScheduler.from_config(M.SD15_SCHEDULER_CONFIG), torch_dtype=dtype)


CASES = {
    "ddim": ("DDIMSchedulerMy", "DDIMScheduler", {}, 6, 7.5),
    "lcm": ("LCMScheduler", "LCMScheduler", {}, 4, 0.0),
    "sde-dpm++": ("DPMSolverScheduler", "DPMSolverScheduler", dict(algorithm_type="sde-dpmsolver++"), 5, 7.5),
    "pndm": ("PNDMScheduler", "PNDMScheduler", {}, 5, 7.5),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rng_only_consumes_what_the_loop_consumes(name, dtype):
    from oracle import schedulers as O
    from oracle.pipeline import denoise, prepare_latents
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    pname, oname, kw, n, g = CASES[name]
    B = 3
    # witness: the oracle loop with a zero UNet draws the initial latents, then the per-step noise
    go = torch.Generator().manual_seed(29)
    pe = torch.zeros(B, 77, 768, dtype=dtype)
    lat = prepare_latents((B, 4, 64, 64), go, "cpu", dtype)
    denoise(_zero_unet, getattr(O, oname).from_config(O.SD15_SCHEDULER_CONFIG, **kw), pe, pe, lat, n, guidance_scale=g,
            generator=go)
    # product: rng_only
    gp = torch.Generator().manual_seed(29)
    pipe = _pipe(dtype=dtype)
    pipe.scheduler = getattr(S, pname).from_config(M.SD15_SCHEDULER_CONFIG, **kw)
    out, secs, x0 = pipe(["p"] * B, num_inference_steps=n, guidance_scale=g, generator=gp, output_type="pt",
                         rng_only=True)
    assert out is None and x0 == []
    assert torch.equal(gp.get_state(), go.get_state()), name
    # and the NEXT batch's latents are therefore identical
    assert torch.equal(torch.randn(4, generator=gp), torch.randn(4, generator=go))


def test_rng_only_two_scheduler_and_skip():
    from oracle import schedulers as O
    from oracle.pipeline import denoise, denoise_two, prepare_latents
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    B, cfg = 2, M.SD15_SCHEDULER_CONFIG
    pe = torch.zeros(B, 77, 768)
    # two-scheduler: no extra kwargs reach the steps (models.py:520 inspects the PNDM default) -> only the latents
    go, gp = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    lat = prepare_latents((B, 4, 64, 64), go, "cpu", torch.float32)
    denoise_two(_zero_unet, O.DDIMScheduler.from_config(cfg), O.DPMSolverScheduler.from_config(cfg), pe, pe, lat, 10, 3)
    pipe = _pipe(M.StableDiffusionModelTwoSchedulers)
    pipe.scheduler_first, pipe.scheduler_second = S.DDIMSchedulerMy.from_config(cfg), S.DPMSolverScheduler.from_config(cfg)
    pipe(["p"] * B, num_inference_steps_first=10, num_inference_steps_second=10, num_step_switch=3, generator=gp,
         output_type="pt", rng_only=True)
    assert torch.equal(gp.get_state(), go.get_state())
    assert pipe.scheduler_first.timesteps.tolist()[:3] == [901, 801, 701]
    # skip-steps with LCM: skipped loop indices draw nothing
    go, gp = torch.Generator().manual_seed(6), torch.Generator().manual_seed(6)
    lat = prepare_latents((B, 4, 64, 64), go, "cpu", torch.float32)
    denoise(_zero_unet, O.LCMScheduler.from_config(cfg), pe, pe, lat, 6, guidance_scale=0, generator=go,
            skip_timesteps=[1, 4])
    pipe = _pipe(M.StableDiffusionModelSkipTimesteps)
    pipe.scheduler = S.LCMScheduler.from_config(cfg)
    pipe(["p"] * B, num_inference_steps=6, guidance_scale=0, generator=gp, output_type="pt", skip_timesteps=[1, 4],
         rng_only=True)
    assert torch.equal(gp.get_state(), go.get_state())


def test_rng_rows_draw_the_global_batch():
    """``rng_rows=(lo, hi, total)``: a rank that computes rows lo:hi of a global batch draws the GLOBAL tensors
    (initial latents and per-step noise) and keeps its rows."""
    from sonicdiffusionbayeslab_b200 import models as M
    from sonicdiffusionbayeslab_b200 import schedulers as S

    total = 8
    g_full, g_shard = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    pipe = _pipe()
    full = pipe.prepare_latents(total, g_full, None, 1.0)
    part = pipe.prepare_latents(3, g_shard, None, 1.0, rng_rows=(2, 5, total))
    assert torch.equal(part, full[2:5])
    s = S.LCMScheduler.from_config(M.SD15_SCHEDULER_CONFIG)
    s.rng_rows = (2, 5, total)
    z_full = S.randn_tensor((total, 4, 64, 64), generator=g_full, device="cpu", dtype=torch.float32)
    z_part = s._draw(torch.empty(3, 4, 64, 64), g_shard, torch.float32)
    assert torch.equal(z_part, z_full[2:5])
    assert torch.equal(g_full.get_state(), g_shard.get_state())


class _FakeModel:
    """Stands in for the pipeline in ``BaseMethod.generate``: every real call draws (B,2) + one extra draw, every
    ``rng_only`` call draws the same and returns nothing."""

    num_timesteps = 1

    def __init__(self):
        self.calls = []

    def __call__(self, prompts, generator=None, rng_only=False, **kw):
        z = torch.randn(len(prompts), 2, generator=generator)
        torch.randn(len(prompts), 2, generator=generator)            # a per-step draw
        self.calls.append(("replay" if rng_only else "run", len(prompts)))
        if rng_only:
            return None, 0.0, []
        from types import SimpleNamespace

        return SimpleNamespace(images=z.reshape(len(prompts), 1, 1, 2)), 0.1, []


@pytest.mark.parametrize("world", [1, 2, 3])
def test_generate_walks_global_batches(world):
    """Every rank of a sharded run walks the global batch list; the union of what the ranks produce equals
    the single-process output, batch for batch (11 prompts, batch 4 -> ragged last batch)."""
    from torch.utils.data import DataLoader, Subset

    from sonicdiffusionbayeslab_b200 import config as cfglib
    from sonicdiffusionbayeslab_b200 import dist as D
    from sonicdiffusionbayeslab_b200.dataset import SyntheticPromptDataset
    from sonicdiffusionbayeslab_b200.experiments.base_experiment import BaseMethod
    from sonicdiffusionbayeslab_b200.metrics.metrics import TimeMetric

    def run(rank, world):
        m = BaseMethod.__new__(type("M", (BaseMethod,), {"run_experiment": lambda self: None}))
        m.config = cfglib.create({"inference": {"batch_size": 4}, "experiment": {"seed": 29}})
        m.rank, m.world = rank, world
        m.test_dataset = SyntheticPromptDataset(n=11, image_size=8)
        m.generator = torch.Generator().manual_seed(29)
        m.model = _FakeModel()
        m.time_metric = TimeMetric()
        imgs, _ = m.generate(m._local_dataloader(4), steps=3, batch_size=4)
        imgs2, _ = m.generate(m._local_dataloader(4), steps=3, batch_size=4)      # next sweep point: same generator
        return imgs + imgs2, m.model.calls, D.shard_batches(11, 4, rank, world)

    single, calls, _ = run(0, 1)
    assert calls == [("run", 4), ("run", 4), ("run", 3)] * 2 and len(single) == 22
    got = [[], []]
    for r in range(world):
        imgs, calls, mine = run(r, world)
        assert len(calls) == 6 and sum(c[0] == "run" for c in calls) == 2 * len(mine)
        n = sum(e - s for s, e in mine)
        got[0] += imgs[:n]
        got[1] += imgs[n:]
    merged = got[0] + got[1]
    assert len(merged) == len(single)
    assert all(torch.equal(a, b) for a, b in zip(merged, single))
