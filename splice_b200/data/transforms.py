"""CPU/PIL augmentation pipeline (drop-in for the reference's data/transforms.py:7-41).

Out of scope as kernels: RNG-order-sensitive host code. It draws from numpy
(`np.random.uniform`, one per call) and torch's CPU generator (torchvision random transforms) in the same
order as the reference, so a given seed yields the same crops.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torchvision.transforms as T


class Global_crops(nn.Module):
    def __init__(self, n_crops, min_cover, last_transform, flip=False):
        super().__init__()
        self.n_crops = n_crops
        self.min_cover = min_cover
        self.last_transform = T.Compose([last_transform] + ([T.RandomHorizontalFlip()] if flip else []))

    def forward(self, img):
        w, h = img.size
        side = int(round(np.random.uniform(self.min_cover * h, h)))
        pipeline = T.Compose([T.RandomCrop(min(side, w)), self.last_transform])
        return torch.stack([pipeline(img) for _ in range(self.n_crops)])


dino_structure_transforms = T.Compose([
    T.RandomHorizontalFlip(p=0.5),
    T.RandomApply([T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.2, hue=0.1)], p=0.5),
    T.RandomApply([T.GaussianBlur(kernel_size=3)], p=0.2),
])

dino_texture_transforms = T.Compose([T.RandomHorizontalFlip(p=0.5)])
