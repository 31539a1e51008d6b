"""Single image-pair dataset (drop-in for the reference's data/Dataset.py:12-73); host-side, unchanged semantics."""
from __future__ import annotations

import os

import torch
from PIL import Image
from torch.utils.data import Dataset
from torchvision import transforms as T

from .transforms import Global_crops, dino_structure_transforms, dino_texture_transforms


class SingleImageDataset(Dataset):
    def __init__(self, cfg):
        self.cfg = cfg
        aug = cfg['use_augmentations']
        self.structure_transforms = dino_structure_transforms if aug else T.Compose([])
        self.texture_transforms = dino_texture_transforms if aug else T.Compose([])
        self.base_transform = T.Compose([T.ToTensor()])
        self.global_A_patches = T.Compose([
            self.structure_transforms,
            Global_crops(n_crops=cfg['global_A_crops_n_crops'], min_cover=cfg['global_A_crops_min_cover'],
                         last_transform=self.base_transform)])
        self.global_B_patches = T.Compose([
            self.texture_transforms,
            Global_crops(n_crops=cfg['global_B_crops_n_crops'], min_cover=cfg['global_B_crops_min_cover'],
                         last_transform=self.base_transform)])

        def first_image(sub):
            d = os.path.join(cfg['dataroot'], sub)
            return Image.open(os.path.join(d, os.listdir(d)[0])).convert('RGB')

        self.A_img, self.B_img = first_image('A'), first_image('B')
        if cfg['A_resize'] > 0:
            self.A_img = T.Resize(cfg['A_resize'])(self.A_img)
        if cfg['B_resize'] > 0:
            self.B_img = T.Resize(cfg['B_resize'])(self.B_img)
        if cfg['direction'] == 'BtoA':
            self.A_img, self.B_img = self.B_img, self.A_img
        print("Image sizes %s and %s" % (str(self.A_img.size), str(self.B_img.size)))
        self.step = torch.zeros(1) - 1

    def get_A(self):
        return self.base_transform(self.A_img).unsqueeze(0)

    def __getitem__(self, index):
        self.step += 1
        sample = {'step': self.step}
        if self.step % self.cfg['entire_A_every'] == 0:
            sample['A'] = self.get_A()
        sample['A_global'] = self.global_A_patches(self.A_img)
        sample['B_global'] = self.global_B_patches(self.B_img)
        return sample

    def __len__(self):
        return 1
