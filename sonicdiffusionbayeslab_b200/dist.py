"""Multi-GPU plumbing: one process per GPU, prompts sharded data-parallel, NCCL only after generation.

The denoising loop shards naturally (one prompt = one trajectory): every rank
holds a full weight replica and runs its own batches with ZERO communication inside the loop.  The only
exchange is after generation, for the CLIP-score metric (/root/reference/calc_clip_score.py,
/root/reference/src/metrics/metrics.py:25-41): an all-gather of the decoded uint8 images and of the
L2-normalised CLIP image/text features, or the cheaper all-reduce of ``(score_sum, n_samples)``
(torchmetrics declares both states ``dist_reduce_fx="sum"``).  Backend: NCCL over NVLink on GPUs, gloo in
the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(device=None):
    """(rank, world); initialises the default process group from torchrun's environment if needed."""
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if world > 1 and not dist.is_initialized():
        # SONIC_DIST_BACKEND=gloo: several ranks on ONE GPU (NCCL refuses duplicate devices) -- used by the tests
        backend = os.environ.get("SONIC_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_batches(n_items: int, batch_size: int, rank: int, world_size: int):
    """Contiguous blocks of WHOLE batches in dataloader order (shuffle=False) -> list of (start, stop)."""
    batches = [(s, min(n_items, s + batch_size)) for s in range(0, n_items, batch_size)]
    per = (len(batches) + world_size - 1) // world_size
    return batches[rank * per:(rank + 1) * per]


def replay_generator_rows(shape, generator, device, dtype, row_start, row_stop):
    """RNG parity with the single-GPU reference (one generator consumed by every batch,
    base_experiment.py:51-53,149): draw the FULL global tensor exactly as one process would, then
    slice this rank's rows -- never draw shard-shaped tensors (Philox mapping depends on the shape)."""
    full = torch.randn(shape, generator=generator, device=device, dtype=dtype)
    return full[row_start:row_stop]


def all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    """Concatenate equally- or unequally-sized row blocks from every rank in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    if dist.get_backend() == "gloo" and t.is_cuda:               # several ranks on one GPU (tests): gather on the host
        return all_gather_cat(t.cpu()).to(t.device)
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.long)
    counts = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in counts]
    dist.all_gather(out, pad.contiguous())
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def quantise_uint8(images: torch.Tensor) -> torch.Tensor:
    """(x * 255).to(uint8): truncation, as base_experiment.py:198-199."""
    return (images * 255).to(torch.uint8)


def gather_images_and_features(images_u8, feat_img, feat_txt):
    """The all-gather of the two_schedulers / calc_clip_score workload: decoded uint8 images
    (786,432 B each) and CLIP feature pairs (2 x 512 fp32 each), in reference order."""
    return all_gather_cat(images_u8), all_gather_cat(feat_img), all_gather_cat(feat_txt)


def clip_score_from_features(feat_img, feat_txt):
    s = 100 * (feat_img * feat_txt).sum(dim=-1)
    return torch.clamp(s.sum() / s.numel(), min=0)


def all_reduce_metric(metric):
    """Sum-reduce a metric's states over ranks (what torchmetrics does at ``compute`` time)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return metric
    t = metric.state_tensor()
    if dist.get_backend() == "gloo" and t.is_cuda:
        t_host = t.cpu()
        dist.all_reduce(t_host, op=dist.ReduceOp.SUM)
        t = t_host.to(t.device)
    elif dist.get_backend() == "nccl" and not t.is_cuda:         # NCCL reduces device tensors only (host-side metrics)
        t_dev = t.to(torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t_dev, op=dist.ReduceOp.SUM)
        t = t_dev.to(t.device)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    metric.load_state_tensor(t)
    return metric
