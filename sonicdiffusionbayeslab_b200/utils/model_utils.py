"""Small helpers of the drivers (/root/reference/src/utils/model_utils.py:15-39)."""
import os
import random

import torch


def setup_seed(seed):
    random.seed(seed)
    torch.random.manual_seed(seed)


def to_pil_image(tensor):
    from torchvision import transforms

    return transforms.ToPILImage()(tensor)


def save_image(image_dir, image_name, image):
    os.makedirs(os.path.join(image_dir, "images"), exist_ok=True)
    image.save(os.path.join(image_dir, "images", f"{image_name.split('.')[0]}.png"))


def save_table(table_dir, table_name, table):
    os.makedirs(table_dir, exist_ok=True)
    table.to_csv(os.path.join(table_dir, f"{table_name}.tsv"), index=None, sep="\t")
