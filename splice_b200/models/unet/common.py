"""Building blocks of the generator (drop-in for the reference's models/unet/common.py:11-124).

What the reference's two skip() call sites need is native (models/networks.py:57 default arguments; inversion.py:21-25
reflection padding, 7x7 / 5x5 filters): zero- or reflection-padded strided convs, BatchNorm2d (always in training
mode), LeakyReLU(0.2), bilinear x2 upsampling, channel concat with centre crop. The DIP leftovers neither call site
reaches (GenNoise, Swish/ELU, lanczos/avg/max down-samplers) are out of scope and raise
NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _append(self, module):
    # the reference names children "1", "2", ... (common.py:5-8); state_dict keys depend on it
    self.add_module(str(len(self) + 1), module)


torch.nn.Module.add = _append


class Concat(nn.Module):
    """Runs every child on the same input and concatenates along `dim`, centre-cropping all results to the
    smallest spatial size (ref common.py:11-42; happens whenever a crop side is odd)."""

    def __init__(self, dim, *args):
        super().__init__()
        self.dim = dim
        for idx, module in enumerate(args):
            self.add_module(str(idx), module)

    def forward(self, input):
        outs = [m(input) for m in self._modules.values()]
        th = min(o.shape[2] for o in outs)
        tw = min(o.shape[3] for o in outs)
        cropped = []
        for o in outs:
            dh, dw = (o.shape[2] - th) // 2, (o.shape[3] - tw) // 2
            cropped.append(o if (dh == 0 and dw == 0 and o.shape[2] == th and o.shape[3] == tw)
                           else o[:, :, dh:dh + th, dw:dw + tw])
        return torch.cat(cropped, dim=self.dim)

    def __len__(self):
        return len(self._modules)


def act(act_fun='LeakyReLU'):
    if isinstance(act_fun, str):
        if act_fun == 'LeakyReLU':
            return nn.LeakyReLU(0.2, inplace=True)
        if act_fun == 'none':
            return nn.Sequential()
        raise NotImplementedError(f"activation {act_fun!r} is outside the splice_b200 hot path")
    return act_fun()


def bn(num_features):
    return nn.BatchNorm2d(num_features)


def conv(in_f, out_f, kernel_size, stride=1, bias=True, pad='zero', downsample_mode='stride'):
    if stride != 1 and downsample_mode != 'stride':
        raise NotImplementedError(f"downsample_mode {downsample_mode!r} is outside the splice_b200 hot path")
    if pad not in ('zero', 'reflection'):
        raise NotImplementedError(f"pad {pad!r} is outside the splice_b200 hot path")
    to_pad = int((kernel_size - 1) / 2)
    if pad == 'reflection':   # ref common.py:113-118: the padder is child "0", the conv child "1" (state_dict keys)
        return nn.Sequential(nn.ReflectionPad2d(to_pad), nn.Conv2d(in_f, out_f, kernel_size, stride, padding=0, bias=bias))
    return nn.Sequential(nn.Conv2d(in_f, out_f, kernel_size, stride, padding=to_pad, bias=bias))
