"""Prompt / image datasets of the experiment drivers (/root/reference/src/dataset/dataset.py:8-41).

``ImageDatasetWithPrompts`` keeps the reference behaviour (``os.listdir`` order, items
``{"image_file", "image", "prompt"}``).  The reference's COCO images are an external download and
its caption JSON is data, not code, so when the configured paths do not exist the drivers fall back
to ``SyntheticPromptDataset``: seeded synthetic captions with the length statistics of the
reference's prompt file (mean ~53 characters) and mid-grey placeholder "real" images.
"""
from __future__ import annotations

import json
import os
import random

import torch
from torch.utils.data import Dataset

_SUBJECTS = ["a man", "a woman", "a dog", "a cat", "two people", "a child", "a group of people", "a bird", "a bus",
             "a train", "a plate of food", "a vase of flowers", "a horse", "a skateboarder", "a tennis player",
             "a laptop", "a kitchen", "a bathroom", "a clock tower", "a parking meter", "a giraffe", "an elephant"]
_VERBS = ["sitting on", "standing next to", "riding", "holding", "looking at", "walking past", "lying on",
          "parked near", "flying over", "eating from", "playing with", "jumping over"]
_OBJECTS = ["a bench", "a table", "the street", "a snowy slope", "a couch", "a field of grass", "a wooden fence",
            "a surfboard", "a red car", "the beach", "a window", "a bed", "a bicycle", "a frisbee", "the sidewalk"]
_TAILS = ["", " in the sun", " at night", " in a city", " near the water", " on a cloudy day", " in black and white"]


def synthetic_prompts(n: int, seed: int = 29) -> list[str]:
    rng = random.Random(seed)
    return [f"{rng.choice(_SUBJECTS)} {rng.choice(_VERBS)} {rng.choice(_OBJECTS)}{rng.choice(_TAILS)}"
            for _ in range(n)]


class SyntheticPromptDataset(Dataset):
    def __init__(self, n: int = 1000, image_size: int = 512, seed: int = 29, prompts=None):
        self.prompts = list(prompts) if prompts is not None else synthetic_prompts(n, seed)
        self.image_size = image_size

    def __len__(self):
        return len(self.prompts)

    def __getitem__(self, idx):
        return {"image_file": f"synthetic_{idx:06d}.png",
                "image": torch.full((3, self.image_size, self.image_size), 0.5),
                "prompt": self.prompts[idx]}


class ImageDatasetWithPrompts(Dataset):
    def __init__(self, image_dir, prompts_file, transform=None):
        from PIL import Image  # noqa: F401

        self.image_dir, self.prompts_file, self.transform = image_dir, prompts_file, transform
        self.image_files = [f for f in os.listdir(image_dir) if os.path.isfile(os.path.join(image_dir, f))]
        with open(prompts_file) as f:
            self.prompts_json = json.load(f)

    def __len__(self):
        return len(self.image_files)

    def __getitem__(self, idx):
        from PIL import Image

        name = self.image_files[idx]
        image = Image.open(os.path.join(self.image_dir, name)).convert("RGB")
        if self.transform:
            image = self.transform(image)
        return {"image_file": name, "image": image, "prompt": self.prompts_json[name]}
