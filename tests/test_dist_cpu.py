"""World-size-2 gloo tests of the multi-GPU host logic (sonicdiffusionbayeslab_b200/dist.py) on CPU.

The denoising loop shards by prompt with no communication; what needs a second rank is the
plumbing around it: contiguous whole-batch sharding in dataloader order, generator replay for RNG parity with the
single-process reference (/root/reference/src/experiments/base_experiment.py:51-53,149), the all-gather of decoded
images + CLIP features and the sum-reduction of metric states (/root/reference/src/metrics/metrics.py:25-41).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from sonicdiffusionbayeslab_b200 import dist as D
    from sonicdiffusionbayeslab_b200.metrics.metrics import TimeMetric

    r, w = D.init_from_env()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    n_items, bs = 23, 4
    # 1. sharding: contiguous whole batches, every item exactly once, in order
    mine = D.shard_batches(n_items, bs, rank, world)
    # 2. RNG parity: every rank replays the ONE shared generator batch by batch and slices its rows
    g = torch.Generator().manual_seed(29)
    all_batches = [(s, min(n_items, s + bs)) for s in range(0, n_items, bs)]
    lat = []
    for (s, e) in all_batches:
        full = D.replay_generator_rows((e - s, 4, 8, 8), g, "cpu", torch.float32, 0, e - s)
        if (s, e) in mine:
            lat.append(full)
    lat = torch.cat(lat) if lat else torch.zeros(0, 4, 8, 8)
    # 3. the "images" of this shard (a deterministic function of the latents) and fake CLIP features
    imgs = D.quantise_uint8(torch.sigmoid(lat.repeat(1, 1, 1, 1)[:, :3]))
    f_img = torch.nn.functional.normalize(lat.flatten(1)[:, :16], dim=-1)
    f_txt = torch.nn.functional.normalize(lat.flatten(1)[:, 16:32], dim=-1)
    gi, gf, gt = D.gather_images_and_features(imgs, f_img, f_txt)     # unequal shard sizes (12 vs 11 items)
    score = D.clip_score_from_features(gf, gt)
    # 4. metric state reduction: seconds / images summed over ranks
    tm = TimeMetric()
    tm.update(torch.tensor(0.5 * (rank + 1)), lat.shape[0])
    D.all_reduce_metric(tm)
    torch.save({"mine": mine, "imgs": gi, "score": score, "time": tm.compute(), "n": gi.shape[0]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_gather_and_metric_reduction(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(world)]
    # shards partition the batch list in order
    batches = [b for o in outs for b in o["mine"]]
    assert batches == [(s, min(23, s + 4)) for s in range(0, 23, 4)]
    # single-process reference: one generator, all batches
    from sonicdiffusionbayeslab_b200 import dist as D

    g = torch.Generator().manual_seed(29)
    lat = torch.cat([torch.randn((e - s, 4, 8, 8), generator=g) for (s, e) in batches])
    imgs = D.quantise_uint8(torch.sigmoid(lat[:, :3]))
    f_img = torch.nn.functional.normalize(lat.flatten(1)[:, :16], dim=-1)
    f_txt = torch.nn.functional.normalize(lat.flatten(1)[:, 16:32], dim=-1)
    ref_score = D.clip_score_from_features(f_img, f_txt)
    for o in outs:                       # every rank holds the full, identically ordered result
        assert o["n"] == 23
        assert torch.equal(o["imgs"], imgs), "gathered images differ from the single-process order"
        assert torch.allclose(o["score"], ref_score, atol=1e-5)
    # seconds summed over ranks / images summed over ranks
    want = (0.5 * 1 + 0.5 * 2) / 23
    assert abs(float(outs[0]["time"]) - want) < 1e-6 and abs(float(outs[1]["time"]) - want) < 1e-6
