"""Device-side sample feed: the reference's per-step augmentation with the image arithmetic on the GPU.

The reference draws every step's sample with PIL on the host (data/Dataset.py:62-70, data/transforms.py:7-41): flip,
colour jitter, blur on the WHOLE structure image, then a random square crop. On the shipped 1200x900 pairs that is
~96 ms per sample (median) against a 5.6 ms optimisation step - no number of host threads on a 16-core box
feeds the loop (16 cores / 96 ms = 166 samples/s with nothing else running).

`DeviceAugmentedDataset` keeps both images resident in HBM and splits every random transform into

  * its PARAMETERS - drawn on the host by torchvision's own `get_params` / `torch.rand(1)` calls, in the reference's
    order, from the same process-wide generators (numpy for the crop side, torch's CPU generator for everything else):
    the random stream is consumed exactly as by `SingleImageDataset` (tests/test_host_logic.py compares the generator
    states after every sample), so a seed selects the same flips, jitter factors, blur sigmas, crop sizes and positions;
  * its ARITHMETIC - torchvision's tensor kernels (`transforms.functional`) applied to the resident fp32 image on a side
    stream. This is a documented NEW pixel stream: PIL truncates to uint8 after every operation (and rotates the hue in an
    8-bit HSV space), the tensor path rounds nowhere. Measured against the PIL pipeline on the same draws: every single
    operation <= 1.1/255 per pixel except the hue rotation (mean 0.9/255, worst pixel 12/255); whole samples: mean
    absolute deviation < 3/255 (PIL's truncation bias adds up over the four jitter steps), worst pixel < 20/255 when a hue
    rotation was drawn, < 3/255 otherwise
    (tests/test_host_logic.py states and checks these bounds). The PIL feed stays the default.

Same interface as `SingleImageDataset` (`dataset[0]` -> sample dict, `get_A()`, `.step`), so it drops into
`PrefetchedSamples` / `train_model` (cfg['device_aug'] = True). Outputs are CUDA tensors that carry their stream's
event as `_splice_ready`, the hand-off `InputStager` uses.
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision.transforms as T
import torchvision.transforms.functional as F

from .Dataset import SingleImageDataset


class DeviceAugmentedDataset:
    def __init__(self, cfg, device=None):
        host = SingleImageDataset(cfg)          # image loading / A_resize / direction: unchanged host logic
        self.cfg = cfg
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.A = T.ToTensor()(host.A_img).to(self.device)      # fp32 [3,H,W] in [0,1], resident
        self.B = T.ToTensor()(host.B_img).to(self.device)
        self.aug = bool(cfg['use_augmentations'])
        self._jitter = T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.2, hue=0.1)    # data/transforms.py:33
        self._blur = T.GaussianBlur(kernel_size=3)                                              # data/transforms.py:36
        self.step = torch.zeros(1) - 1
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._alive = []                          # recent samples: their memory must outlive the consumer's kernels

    # ---- the reference's transforms, parameters on the host / arithmetic on the device ----------------
    def _structure(self, img):
        """dino_structure_transforms (data/transforms.py:30-37)"""
        if torch.rand(1) < 0.5:                                    # RandomHorizontalFlip(p=0.5).forward
            img = F.hflip(img)
        if not (0.5 < torch.rand(1)):                              # RandomApply(p=0.5).forward: skipped when p < rand
            j = self._jitter
            order, b, c, s, h = T.ColorJitter.get_params(j.brightness, j.contrast, j.saturation, j.hue)
            for fn_id in order:                                    # ColorJitter.forward
                if fn_id == 0 and b is not None:
                    img = F.adjust_brightness(img, b)
                elif fn_id == 1 and c is not None:
                    img = F.adjust_contrast(img, c)
                elif fn_id == 2 and s is not None:
                    img = F.adjust_saturation(img, s)
                elif fn_id == 3 and h is not None:
                    img = F.adjust_hue(img, h)
        if not (0.2 < torch.rand(1)):                              # RandomApply([GaussianBlur(3)], p=0.2)
            sigma = T.GaussianBlur.get_params(self._blur.sigma[0], self._blur.sigma[1])
            img = F.gaussian_blur(img, list(self._blur.kernel_size), [sigma, sigma])
        return img

    @staticmethod
    def _texture(img):
        """dino_texture_transforms (data/transforms.py:39-41)"""
        if torch.rand(1) < 0.5:
            img = F.hflip(img)
        return img

    @staticmethod
    def _global_crops(img, n_crops, min_cover):
        """Global_crops.forward (data/transforms.py:20-28): one side per call (numpy), one position per crop (torch)"""
        h, w = img.shape[-2], img.shape[-1]
        side = int(round(np.random.uniform(min_cover * h, h)))
        side = min(side, w)
        crops = []
        for _ in range(n_crops):
            i, j, th, tw = T.RandomCrop.get_params(img, (side, side))
            crops.append(img[:, i:i + th, j:j + tw])
        return torch.stack(crops).contiguous()

    def get_A(self):
        return self.A.unsqueeze(0)

    def __getitem__(self, index):
        self.step += 1
        sample = {'step': self.step}
        cfg = self.cfg
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            if self.step % cfg['entire_A_every'] == 0:
                sample['A'] = self.get_A()
            a = self._structure(self.A) if self.aug else self.A
            sample['A_global'] = self._global_crops(a, cfg['global_A_crops_n_crops'], cfg['global_A_crops_min_cover'])
            b = self._texture(self.B) if self.aug else self.B
            sample['B_global'] = self._global_crops(b, cfg['global_B_crops_n_crops'], cfg['global_B_crops_min_cover'])
            if self.stream is not None:
                ev = torch.cuda.Event()
                ev.record(self.stream)
                for k, v in sample.items():
                    if k != 'step':
                        v._splice_ready = ev
        self._alive.append(sample)
        if len(self._alive) > 32:
            self._alive.pop(0)
        return sample

    def __len__(self):
        return 1


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
