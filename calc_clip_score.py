"""CLIP score of a folder of images against their prompts -- the B200 counterpart of
/root/reference/calc_clip_score.py:13-97 (same command line: ``--folder_path --prompts_file --batch_size
--model_name_or_path``).

    python calc_clip_score.py --folder_path ./data/generate_images/... --prompts_file ./data/dataset/img2annotations_test.json
    python -m torch.distributed.run --nproc-per-node 8 calc_clip_score.py ...        # images sharded over the GPUs

Differences from the reference, each forced by a defect or by the deployment:
  * images go to the metric as uint8 (``pil_to_tensor``), the form torchmetrics' CLIPScore expects and the
    in-pipeline path uses (base_experiment.py:198-201); the reference feeds float [0,1] tensors from ``ToTensor``
    (calc_clip_score.py:69-73), which the HF processor rescales a second time;
    ``--literal_float_rescale`` reproduces that arithmetic for a like-for-like number;
  * preprocessing (bicubic + antialias resize, centre crop, normalise) and both CLIP towers run on the GPU
    (metrics/metrics.py, clip_engine.py) instead of PIL on the host + library modules;
  * under torchrun the image list is sharded in contiguous blocks and the ONLY collective of the whole system runs
    here: an NCCL all-gather of the CLIP image / text features (and, with ``--gather_images``, of the uint8 images),
    after which every rank holds the features in dataset order and computes the same score;
  * ``--synthetic N`` scores N seeded synthetic images/captions (no dataset exists offline).
"""
import argparse
import os

import torch

from sonicdiffusionbayeslab_b200 import dist as D
from sonicdiffusionbayeslab_b200.dataset import ImageDatasetWithPrompts
from sonicdiffusionbayeslab_b200.dataset.dataset import synthetic_prompts
from sonicdiffusionbayeslab_b200.metrics.metrics import ClipScoreMetric


class _SyntheticImages(torch.utils.data.Dataset):
    """Seeded uint8 noise images with seeded captions: the offline stand-in for a folder of generated images."""

    def __init__(self, n, size=512, seed=29):
        self.prompts = synthetic_prompts(n, seed)
        self.size, self.seed = size, seed

    def __len__(self):
        return len(self.prompts)

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        img = torch.randint(0, 256, (3, self.size, self.size), generator=g, dtype=torch.uint8)
        return {"image_file": f"synthetic_{i:05d}.png", "image": img, "prompt": self.prompts[i]}


def _to_uint8(images):
    if images.dtype == torch.uint8:
        return images
    return (images * 255).to(torch.uint8)              # ToTensor() floats in [0, 1] -> the uint8 they came from


def calc_clip_score(dataset, model_name_or_path="openai/clip-vit-base-patch16", device=None, batch_size=32,
                    gather_images=False, literal_float_rescale=False):
    """Returns (score, n_images[, gathered uint8 images]).  Sharded over the default process group if one exists.
    ``literal_float_rescale``: score what the reference script literally scores -- its float [0,1] tensors are
    rescaled by 1/255 a second time inside the HF processor (SURVEY C-8: near-black images)."""
    rank, world = D.world()
    if device is None:
        device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0)) % max(1, torch.cuda.device_count())}"
    metric = ClipScoreMetric(model_name_or_path=model_name_or_path).to(device)
    metric.literal_float_rescale = bool(literal_float_rescale)
    n = len(dataset)
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    loader = torch.utils.data.DataLoader(torch.utils.data.Subset(dataset, range(lo, hi)), batch_size=batch_size,
                                         shuffle=False)
    fi, ft, imgs = [], [], []
    for batch in loader:
        u8 = _to_uint8(batch["image"]).to(device)
        a, b = metric.features(u8, list(batch["prompt"]))
        fi.append(a)
        ft.append(b)
        if gather_images:
            imgs.append(u8)
    empty = torch.zeros(0, 512, device=device)
    fi, ft = (torch.cat(fi) if fi else empty), (torch.cat(ft) if ft else empty)
    gi = D.all_gather_cat(torch.cat(imgs)) if gather_images and imgs else None
    fi, ft = D.all_gather_cat(fi), D.all_gather_cat(ft)                     # NCCL over NVLink
    score = D.clip_score_from_features(fi, ft).item()
    return (score, fi.shape[0], gi) if gather_images else (score, fi.shape[0])


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Calculate CLIP score for images and prompts")
    parser.add_argument("--folder_path", type=str, help="Path to folder containing images")
    parser.add_argument("--prompts_file", type=str, help="JSON file containing prompts for images")
    parser.add_argument("--batch_size", type=int, default=32, help="Batch size for processing")
    parser.add_argument("--model_name_or_path", type=str, default="openai/clip-vit-base-patch16",
                        help="CLIP model name or path to use for scoring")
    parser.add_argument("--synthetic", type=int, default=0, help="score N seeded synthetic images instead of a folder")
    parser.add_argument("--gather_images", action="store_true", help="also all-gather the uint8 images")
    parser.add_argument("--literal_float_rescale", action="store_true",
                        help="reproduce the reference script's literal arithmetic: float [0,1] inputs rescaled by 1/255 "
                             "a second time by the HF processor (calc_clip_score.py:68-72, SURVEY C-8)")
    args = parser.parse_args()

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) % max(1, torch.cuda.device_count()))
    torch.cuda.set_device(dev)
    rank, world = D.init_from_env(dev)
    if args.synthetic:
        dataset = _SyntheticImages(args.synthetic)
    else:
        if not args.folder_path or not os.path.isdir(args.folder_path):
            raise ValueError("Please provide a valid folder path containing images")
        if not args.prompts_file or not os.path.isfile(args.prompts_file):
            raise ValueError("Please provide a valid JSON file containing prompts")
        from torchvision import transforms

        dataset = ImageDatasetWithPrompts(image_dir=args.folder_path, prompts_file=args.prompts_file,
                                          transform=transforms.Compose([transforms.PILToTensor()]))
    out = calc_clip_score(dataset, model_name_or_path=args.model_name_or_path, device=dev, batch_size=args.batch_size,
                          gather_images=args.gather_images, literal_float_rescale=args.literal_float_rescale)
    if rank == 0:
        print(f"CLIP Score: {out[0]}")
    if world > 1:
        torch.distributed.destroy_process_group()
