"""Metric plugins (/root/reference/src/metrics/metrics.py): ``clip_score``, ``time_metric``, and
``image_reward`` / ``fid`` placeholders.

torchmetrics is not installable here, so a minimal ``Metric`` base provides the surface the drivers
use (``update`` / ``compute`` / ``reset`` / ``to``; sum-reduced states).  ``ClipScoreMetric`` restates
torchmetrics 1.6.1 ``CLIPScore`` on ``transformers.CLIPModel``:
``max(0, mean_i 100 * cos(f_img(i), f_txt(i)))`` -- with the image preprocessing (resize shortest side
to 224 with PIL's antialiased bicubic resampler, centre crop, 1/255, CLIP mean/std) done by ONE native kernel
(csrc/preprocess.cu, bit-identical to PIL) instead of PIL on the host, which is where the reference spends its
metric time, and both towers on the native engine (clip_engine.py).
ImageReward and FID need downloaded networks and the real COCO images (no network): they stay
registered so ``BaseMethod.setup_metrics`` works, and report NaN.
"""
from __future__ import annotations

import logging
import os
import warnings

import torch

from ..registry import metrics_registry
from ..text import HashTokenizer, load_tokenizer

from ..kernels import CLIP_MEAN, CLIP_STD  # noqa: E402,F401


class Metric:
    """Sum-reduced metric states, torchmetrics-style."""

    def __init__(self):
        self._defaults = {}
        self.device = torch.device("cpu")

    def add_state(self, name, default, dist_reduce_fx="sum"):
        self._defaults[name] = default.clone()
        setattr(self, name, default.clone())

    def reset(self):
        for k, v in self._defaults.items():
            setattr(self, k, v.clone().to(self.device))

    def to(self, device):
        self.device = torch.device(device)
        for k in self._defaults:
            setattr(self, k, getattr(self, k).to(self.device))
        return self

    def state_tensor(self):
        return torch.stack([getattr(self, k).double().to(self.device) for k in self._defaults])

    def load_state_tensor(self, t):
        for k, v in zip(self._defaults, t):
            setattr(self, k, v.to(getattr(self, k).dtype))


def clip_preprocess(images: torch.Tensor, size: int = 224) -> torch.Tensor:
    """uint8 (n,3,H,W) -> normalised float (n,3,size,size): HF ``CLIPImageProcessor`` semantics (PIL bicubic with
    antialiasing, centre crop, 1/255, CLIP mean/std) in ONE native kernel, bit-identical to PIL's uint8 resize
    (``kernels.clip_preprocess`` / csrc/preprocess.cu)."""
    if images.dtype != torch.uint8:
        raise TypeError("CLIP score expects uint8 images (the in-pipeline path, base_experiment.py:198-201); "
                        "float [0,1] inputs are the calc_clip_score.py defect (SURVEY C-8)")
    from .. import kernels as K

    return K.clip_preprocess(images.contiguous(), size=size)


class ClipWeights:
    """State dict + config of a ``transformers.CLIPModel`` (ViT-B/16): the towers run on the native engine
    (clip_engine.py), the module that produced the weights is discarded."""

    def __init__(self, state_dict, config, source):
        self._sd, self.config, self.source = state_dict, config, source

    def state_dict(self):
        return self._sd


def make_clip_model(model_name_or_path=None, seed=29):
    """CLIP ViT-B/16 weights from a local directory, else seeded random-init of that architecture (with a loud
    warning: the score of a random-init CLIP is a synthetic-workload number).  Returns (ClipWeights, tokenizer)."""
    from transformers import CLIPConfig, CLIPModel

    if model_name_or_path and os.path.isdir(model_name_or_path):
        model, tok, source = CLIPModel.from_pretrained(model_name_or_path).eval(), load_tokenizer(model_name_or_path), \
            "provided"
    else:
        msg = (f"clip_score: no local weights at {model_name_or_path!r} and no network -- RANDOM-INIT CLIP ViT-B/16 and a "
               "CRC32 hash tokenizer are used; the reported CLIP score is a synthetic-workload number")
        if os.environ.get("SONIC_REQUIRE_WEIGHTS") == "1":
            raise FileNotFoundError(msg)
        logging.getLogger("sonicdiffusionbayeslab_b200").warning(msg)
        warnings.warn(msg, RuntimeWarning, stacklevel=2)
        cfg = CLIPConfig(text_config=dict(vocab_size=49408, hidden_size=512, intermediate_size=2048,
                                          num_hidden_layers=12, num_attention_heads=8, max_position_embeddings=77,
                                          bos_token_id=49406, eos_token_id=49407, pad_token_id=1),
                         vision_config=dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                            num_attention_heads=12, image_size=224, patch_size=16),
                         projection_dim=512)
        state = torch.random.get_rng_state()
        torch.manual_seed(seed + 3)
        try:
            model = CLIPModel(cfg)
        finally:
            torch.random.set_rng_state(state)
        tok, source = HashTokenizer(), "random-init"
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return ClipWeights(sd, model.config, source), tok


@metrics_registry.add_to_registry("clip_score")
class ClipScoreMetric(Metric):
    def __init__(self, model_name_or_path: str = "openai/clip-vit-base-patch16", **kwargs):
        super().__init__()
        self.model, self.tokenizer = make_clip_model(model_name_or_path)
        self.add_state("score", torch.tensor(0.0))
        self.add_state("n_samples", torch.tensor(0, dtype=torch.long))
        self.keep_features = False
        self._feats = []
        self._engines = {}

    @torch.no_grad()
    def features(self, images, text):
        """L2-normalised (image, text) features of the native towers; uint8 images (n,3,H,W)."""
        if self.device.type != "cuda":
            raise RuntimeError("clip_score runs on the B200 engine (clip_engine.py): call .to('cuda') first")
        if isinstance(images, (list, tuple)):
            images = torch.stack(list(images))
        if images.dim() == 3:
            images = images[None]
        text = [text] if isinstance(text, str) else list(text)
        if len(text) != images.shape[0]:
            raise ValueError("Expected the number of images and text examples to be the same")
        if images.dtype != torch.uint8:
            raise TypeError("CLIP score expects uint8 images (the in-pipeline path, base_experiment.py:198-201); "
                            "float [0,1] inputs are the calc_clip_score.py defect (SURVEY C-8)")
        ids, _ = self.tokenizer(text)
        vis, txt = self._native(len(text))
        fi = vis.image_features_from_images(images.to(self.device),             # fused preprocess -> patch rows
                                            rescale_twice=getattr(self, "literal_float_rescale", False)).float()
        ft = txt.text_features(ids.to(self.device)).float()
        fi = fi / fi.norm(p=2, dim=-1, keepdim=True)
        ft = ft / ft.norm(p=2, dim=-1, keepdim=True)
        return fi, ft

    def _native(self, n):
        """Native towers for a batch of n (built on first use from the transformers-layout state dict)."""
        if n not in self._engines:
            from ..clip_engine import ClipTextEngine, ClipVisionEngine

            sd = self.model.state_dict()
            vc, tc = self.model.config.vision_config, self.model.config.text_config
            self._engines[n] = (
                ClipVisionEngine(sd, n=n, image_size=vc.image_size, patch=vc.patch_size, width=vc.hidden_size,
                                 heads=vc.num_attention_heads, layers=vc.num_hidden_layers, mlp=vc.intermediate_size,
                                 device=self.device),
                ClipTextEngine(sd, n=n, seq=tc.max_position_embeddings, width=tc.hidden_size,
                               heads=tc.num_attention_heads, layers=tc.num_hidden_layers, mlp=tc.intermediate_size,
                               device=self.device))
        return self._engines[n]

    def update(self, images, text):
        fi, ft = self.features(images, text)
        score = 100 * (fi * ft).sum(dim=-1)
        self.score = self.score.to(self.device) + score.sum()
        self.n_samples = self.n_samples.to(self.device) + len(score)
        if self.keep_features:
            self._feats.append((fi, ft))

    def compute(self):
        return torch.max(self.score / self.n_samples, torch.zeros_like(self.score))

    def reset(self):
        super().reset()
        self._feats = []

    def calc_metric(self, data, prompts, batch_size: int = 4) -> float:
        """metrics.py:27-41: score a list of PIL images against prompts."""
        from torchvision.transforms.functional import pil_to_tensor

        tensors = [pil_to_tensor(img) for img in data]
        for i in range(0, len(tensors), batch_size):
            self.update(torch.stack(tensors[i:i + batch_size]), list(prompts[i:i + batch_size]))
        return self.compute().item()


@metrics_registry.add_to_registry("time_metric")
class TimeMetric(Metric):
    """metrics.py:115-131: seconds of denoising-loop time per image."""

    def __init__(self):
        super().__init__()
        self.add_state("time", torch.tensor(0.0))
        self.add_state("total", torch.tensor(0))

    def update(self, time: float, batch_size: int) -> None:
        self.time = self.time + torch.tensor(float(time))
        self.total = self.total + torch.tensor(int(batch_size))

    def compute(self):
        return self.time / self.total


class _Unavailable(Metric):
    reason = ""

    def update(self, *a, **k):
        pass

    def compute(self):
        return torch.tensor(float("nan"))


@metrics_registry.add_to_registry("image_reward")
class RewardModel(_Unavailable):
    reason = "ImageReward-v1.0 weights are an external download"

    def __init__(self, model_name: str = "ImageReward-v1.0", device: str = "cpu", **kwargs):
        super().__init__()


@metrics_registry.add_to_registry("fid")
class FID(_Unavailable):
    reason = "Inception weights and the real COCO images are external downloads"

    def __init__(self, feature: int = 64, input_img_size: int = 512, normalize: bool = False, **kwargs):
        super().__init__()
