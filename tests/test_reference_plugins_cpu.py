"""Metric plugins, dataset and driver helpers against the REFERENCE'S OWN SOURCE, executed where it lies
(oracle/refexec.py ``load_plugins``): /root/reference/src/metrics/metrics.py (``TimeMetric`` :115-131 -- SURVEY 8 row
a13 --, ``ClipScoreMetric.calc_metric`` :25-41 -- row a12's accumulation), src/dataset/dataset.py,
src/utils/model_utils.py.  The CLIP towers are replaced on BOTH sides by the same deterministic feature function, so
what is compared is the reference's batching and the metric-state arithmetic (the towers and the preprocessing have
their own parity tests against ``transformers`` / PIL on the GPU).  Skipped where the reference tree is absent.
"""
import json
import os
import random
import sys
import zlib

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refexec  # noqa: E402

pytestmark = pytest.mark.skipif(not refexec.available(), reason="/root/reference is absent (GPU box)")


def fake_features(images, text):
    """Deterministic stand-in for the two CLIP towers: unit vectors from image statistics and a text hash; about
    half of the pairs have a negative cosine."""
    if isinstance(images, (list, tuple)):
        images = torch.stack(list(images))
    if images.dim() == 3:
        images = images[None]
    text = [text] if isinstance(text, str) else list(text)
    x = images.float()
    fi = torch.stack([x.mean(dim=(1, 2, 3)) - 127.0, x[:, 0].std(dim=(1, 2)), x[:, 1, ::2].mean(dim=(1, 2)) - 120.0,
                      x[:, 2, :, ::3].amax(dim=(1, 2)) - 250.0], dim=1)
    ft = torch.tensor([[((zlib.crc32(f"{k}:{t}".encode()) % 2001) - 1000) / 1000.0 for k in range(4)] for t in text])
    return fi / fi.norm(dim=-1, keepdim=True), ft / ft.norm(dim=-1, keepdim=True)


@pytest.fixture(scope="module")
def ref():
    return refexec.load_plugins(features=fake_features)


@pytest.fixture(autouse=True)
def _one_random_init_clip(monkeypatch):
    """``ClipScoreMetric()`` builds a random-init CLIP ViT-B/16 (seconds each); the towers are replaced by
    ``fake_features`` in every test here, so one instance of the weights serves them all -- the warning still fires."""
    import warnings

    from sonicdiffusionbayeslab_b200.metrics import metrics as MM

    real = MM.make_clip_model

    def cached(model_name_or_path=None, seed=29):
        if "clip" not in _CACHE:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _CACHE["clip"] = real(model_name_or_path, seed)
        warnings.warn("clip_score: RANDOM-INIT CLIP (cached for the tests)", RuntimeWarning, stacklevel=2)
        return _CACHE["clip"]

    monkeypatch.setattr(MM, "make_clip_model", cached)


_CACHE = {}


def test_time_metric_equals_reference_source(ref):
    from sonicdiffusionbayeslab_b200.metrics.metrics import TimeMetric

    rng = random.Random(5)
    mine, theirs = TimeMetric(), ref.metrics.TimeMetric()
    for round_ in range(3):                                   # reset between sweep points (base_experiment.py:250-256)
        for _ in range(rng.randint(1, 7)):
            secs, bs = rng.uniform(0.01, 40.0), rng.choice([1, 2, 16, 32])
            mine.update(secs, bs)
            theirs.update(secs, bs)
            a, b = mine.compute(), theirs.compute()
            assert a.dtype == b.dtype == torch.float32 and torch.equal(a, b), (round_, a, b)
        mine.reset()
        theirs.reset()
    # the last, partial batch is still counted as ``batch_size`` images (base_experiment.py:161): the metric only
    # ever sees what the driver passes, so 10 prompts in batches of 4 report time / 12
    for m in (mine, theirs):
        for secs in (2.0, 2.0, 1.0):
            m.update(secs, 4)
    assert torch.equal(mine.compute(), theirs.compute()) and abs(mine.compute().item() - 5.0 / 12) < 1e-7


def test_metric_registry_names_equal_reference_source(ref):
    from sonicdiffusionbayeslab_b200.registry import metrics_registry

    assert sorted(metrics_registry.classes) == sorted(ref.registry.metrics_registry.classes) == [
        "clip_score", "fid", "image_reward", "time_metric"]


@pytest.mark.parametrize("n,batch_size", [(10, 4), (8, 4), (3, 32)])
def test_clip_score_accumulation_equals_reference_source(ref, n, batch_size, monkeypatch):
    """``calc_metric`` (metrics.py:27-41): pil_to_tensor, batches of ``batch_size`` with a ragged tail, ``update`` per
    batch, ``compute().item()``; state arithmetic of torchmetrics ``CLIPScore`` (sum of 100 cos, count, max(., 0))."""
    from PIL import Image

    from sonicdiffusionbayeslab_b200.metrics.metrics import ClipScoreMetric

    g = np.random.default_rng(n * 100 + batch_size)
    images = [Image.fromarray(g.integers(0, 256, (24, 24, 3), dtype=np.uint8)) for _ in range(n)]
    prompts = [f"caption number {i} of {n}" for i in range(n)]
    theirs = ref.metrics.ClipScoreMetric(model_name_or_path="openai/clip-vit-base-patch16")
    want = theirs.calc_metric(images, prompts, batch_size=batch_size)

    seen = []

    def features(self, imgs, text):
        seen.append((imgs, list(text)))
        return fake_features(imgs, text)

    monkeypatch.setattr(ClipScoreMetric, "features", features)
    with pytest.warns(RuntimeWarning, match="RANDOM-INIT CLIP"):
        mine = ClipScoreMetric(model_name_or_path="openai/clip-vit-base-patch16")
    got = mine.calc_metric(images, prompts, batch_size=batch_size)
    assert got == want                                                    # same fp32 accumulation order
    assert len(seen) == len(theirs.seen) == -(-n // batch_size)
    for (ia, ta), (ib, tb) in zip(seen, theirs.seen):
        assert torch.equal(ia, ib) and ia.dtype == torch.uint8 and ta == tb
    assert int(mine.n_samples) == int(theirs.n_samples) == n
    # a negative mean is clamped to 0 by compute() on both sides
    mine.reset()
    theirs.reset()
    neg = lambda imgs, text: (torch.tensor([[1.0, 0.0]]).repeat(len(text), 1),          # noqa: E731
                              torch.tensor([[-1.0, 0.0]]).repeat(len(text), 1))
    monkeypatch.setattr(ClipScoreMetric, "features", lambda self, i, t: neg(i, t))
    mine.update(torch.zeros(2, 3, 8, 8, dtype=torch.uint8), ["a", "b"])
    assert mine.compute().item() == 0.0 and float(mine.score) == -200.0


def test_dataset_items_equal_reference_source(ref, tmp_path):
    from PIL import Image
    from torchvision import transforms

    from sonicdiffusionbayeslab_b200.dataset.dataset import ImageDatasetWithPrompts

    g = np.random.default_rng(0)
    prompts = {}
    for i in range(6):
        name = f"img_{(i * 7) % 6}.{'png' if i % 2 else 'jpg'}"
        Image.fromarray(g.integers(0, 256, (16 + i, 20, 3), dtype=np.uint8)).convert("L" if i == 3 else "RGB").save(
            tmp_path / name)
        prompts[name] = f"prompt for {name}"
    (tmp_path / "a_directory").mkdir()                                    # not a file: skipped by both
    pj = tmp_path.parent / f"{tmp_path.name}_prompts.json"
    pj.write_text(json.dumps(prompts))
    for tf in (None, transforms.Compose([transforms.PILToTensor()])):
        a = ImageDatasetWithPrompts(str(tmp_path), str(pj), transform=tf)
        b = ref.dataset.ImageDatasetWithPrompts(str(tmp_path), str(pj), transform=tf)
        assert len(a) == len(b) == 6 and a.image_files == b.image_files
        for i in range(len(a)):
            x, y = a[i], b[i]
            assert x.keys() == y.keys() and x["image_file"] == y["image_file"] and x["prompt"] == y["prompt"]
            if tf is None:
                assert x["image"].mode == y["image"].mode == "RGB" and x["image"].tobytes() == y["image"].tobytes()
            else:
                assert torch.equal(x["image"], y["image"])


def test_driver_helpers_equal_reference_source(ref, tmp_path):
    import pandas as pd
    from PIL import Image

    from sonicdiffusionbayeslab_b200.utils import model_utils as mine

    theirs = ref.model_utils
    draws = []
    for mod in (mine, theirs):
        mod.setup_seed(29)
        draws.append((random.random(), torch.rand(3).tolist()))
    assert draws[0] == draws[1]
    t = torch.rand(3, 9, 7)
    assert mine.to_pil_image(t).tobytes() == theirs.to_pil_image(t).tobytes()
    table = pd.DataFrame({"nfe": [3, 5], "clip_score_gen_image": [21.5, 22.25], "time_metric": [0.1, 0.2]})
    mine.save_table(str(tmp_path / "a"), "metrics", table)
    theirs.save_table(str(tmp_path / "b"), "metrics", table)
    assert (tmp_path / "a" / "metrics.tsv").read_bytes() == (tmp_path / "b" / "metrics.tsv").read_bytes()
    img = Image.fromarray(np.arange(48, dtype=np.uint8).reshape(4, 4, 3))
    mine.save_image(str(tmp_path / "ia"), "000123.jpg", img)
    theirs.save_image(str(tmp_path / "ib"), "000123.jpg", img)            # fresh directory: the reference's happy path
    assert (tmp_path / "ia" / "images" / "000123.png").read_bytes() == (tmp_path / "ib" / "images" / "000123.png").read_bytes()
    # the reference only creates ``images/`` when ``image_dir`` itself is missing (model_utils.py:26-28) and fails on
    # an existing directory without it; the product creates it either way
    (tmp_path / "ic").mkdir()
    mine.save_image(str(tmp_path / "ic"), "x.png", img)
    (tmp_path / "id").mkdir()
    with pytest.raises(FileNotFoundError):
        theirs.save_image(str(tmp_path / "id"), "x.png", img)


@pytest.mark.parametrize("n,batch_size", [(7, 3), (4, 32)])
def test_calc_clip_score_function_equals_reference_source(ref, tmp_path, n, batch_size, monkeypatch):
    """/root/reference/calc_clip_score.py:13-37 (one CLIPScore fed batch by batch) against the product's root script
    function (features per batch, one score from all features -- the form that all-gathers under sharding): same
    images and prompts in, same score out up to the order of the fp32 sum."""
    import importlib.util

    from PIL import Image
    from torchvision import transforms

    from sonicdiffusionbayeslab_b200.metrics.metrics import ClipScoreMetric

    g = np.random.default_rng(n)
    prompts = {}
    for i in range(n):
        Image.fromarray(g.integers(0, 256, (20, 20, 3), dtype=np.uint8)).save(tmp_path / f"{i:03d}.png")
        prompts[f"{i:03d}.png"] = f"a synthetic caption {i}"
    pj = tmp_path.parent / f"{tmp_path.name}_prompts.json"
    pj.write_text(json.dumps(prompts))
    tf = transforms.Compose([transforms.PILToTensor()])                   # uint8, the form CLIPScore expects (C-8)
    ds = ref.dataset.ImageDatasetWithPrompts(str(tmp_path), str(pj), transform=tf)
    want = ref.calc_clip_score(torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False),
                               model_name_or_path="openai/clip-vit-base-patch16", device="cpu", batch_size=batch_size)

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("product_calc_clip_score", os.path.join(root, "calc_clip_score.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(ClipScoreMetric, "features", lambda self, imgs, text: fake_features(imgs, text))
    with pytest.warns(RuntimeWarning, match="RANDOM-INIT CLIP"):
        got, count = mod.calc_clip_score(mod.ImageDatasetWithPrompts(str(tmp_path), str(pj), transform=tf),
                                         model_name_or_path="openai/clip-vit-base-patch16", device="cpu",
                                         batch_size=batch_size)
    assert count == n
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)), (got, want)
