"""NCCL sanity of the multi-GPU path on real GPUs (run under torchrun, >= 2 ranks):
prompts sharded in whole batches, generator replay for RNG parity, the engine on every rank, then the all-gather of
(quantised) outputs + CLIP-like features and the metric all-reduce -- compared on rank 0 with a single-process run
over all prompts.  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_nccl_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from sonicdiffusionbayeslab_b200 import dist as D
from sonicdiffusionbayeslab_b200 import models as M
from sonicdiffusionbayeslab_b200 import schedulers as S
from sonicdiffusionbayeslab_b200.metrics.metrics import TimeMetric
from sonicdiffusionbayeslab_b200.text import HashTokenizer
from sonicdiffusionbayeslab_b200.unet_spec import random_unet_state_dict

local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
rank, world = D.init_from_env(dev)
assert world >= 2 and dist.get_backend() == "nccl"

n_items, bs, steps = 10, 4, 2
g = torch.Generator().manual_seed(7)
pe_all = torch.randn(n_items, 77, 768, generator=g)
ne_all = torch.randn(n_items, 77, 768, generator=g)


def make():
    m = M.StableDiffusionModel(random_unet_state_dict(29), vae=None, text_encoder=None, tokenizer=HashTokenizer(),
                               scheduler=S.DDIMSchedulerMy.from_config(M.SD15_SCHEDULER_CONFIG), torch_dtype=torch.bfloat16)
    m.device = dev
    return m


def run(model, batches, mine):
    gen = torch.Generator(device=dev).manual_seed(29)            # ONE generator, consumed batch by batch by everyone
    outs = []
    for (s, e) in batches:
        lat = D.replay_generator_rows((e - s, 4, 64, 64), gen, dev, torch.bfloat16, 0, e - s)
        if (s, e) in mine:
            o, _, _ = model(prompt_embeds=pe_all[s:e].to(dev), negative_prompt_embeds=ne_all[s:e].to(dev), latents=lat,
                            num_inference_steps=steps, guidance_scale=7.5, output_type="latent")
            outs.append(o.images.float())
    return torch.cat(outs) if outs else torch.zeros(0, 4, 64, 64, device=dev)


model = make()
batches = [(s, min(n_items, s + bs)) for s in range(0, n_items, bs)]
mine = D.shard_batches(n_items, bs, rank, world)
lat = run(model, batches, mine)
imgs = D.quantise_uint8(torch.sigmoid(lat[:, :3]))
f_img = torch.nn.functional.normalize(lat.flatten(1)[:, :512], dim=-1)
f_txt = torch.nn.functional.normalize(lat.flatten(1)[:, 512:1024], dim=-1)
gi, gf, gt = D.gather_images_and_features(imgs, f_img, f_txt)    # NCCL all-gather, unequal shard sizes
score = D.clip_score_from_features(gf, gt)
tm = TimeMetric()
tm.update(torch.tensor(0.25 * (rank + 1), device=dev), lat.shape[0])
D.all_reduce_metric(tm)                                          # NCCL all-reduce of (seconds, images)
if rank == 0:
    ref = run(model, batches, batches)                           # single process, every batch
    ref_imgs = D.quantise_uint8(torch.sigmoid(ref[:, :3]))
    r_img = torch.nn.functional.normalize(ref.flatten(1)[:, :512], dim=-1)
    r_txt = torch.nn.functional.normalize(ref.flatten(1)[:, 512:1024], dim=-1)
    assert gi.shape[0] == n_items and torch.equal(gi, ref_imgs), "gathered outputs differ from the single-process run"
    assert torch.allclose(score, D.clip_score_from_features(r_img, r_txt), atol=1e-4)
    want = sum(0.25 * (r + 1) for r in range(world)) / n_items
    assert abs(float(tm.compute()) - want) < 1e-6
    print(f"nccl check ok: world {world}, {n_items} items in shards {[len(D.shard_batches(n_items, bs, r, world)) for r in range(world)]} "
          f"batches, gathered outputs bit-identical to the single-process run, score {float(score):.4f}", flush=True)
dist.barrier()
dist.destroy_process_group()
