"""Generator topology (drop-in for the reference's models/unet/skip.py:4-102).

Builds the same nested nn.Sequential tree — hence the same `state_dict()` keys, `parameters()` order and,
under the same seed, the same initial weights — by recursion over the scales instead of the reference's
iterative pointer-chasing construction.
"""
from __future__ import annotations

import torch.nn as nn

from ...generator import NativeSkip
from ...generator_x import NativeSkipX
from .common import Concat, act, bn, conv


def _listify(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


def skip(
        num_input_channels=3, num_output_channels=3,
        num_channels_down=[16, 32, 64, 128, 128], num_channels_up=[16, 32, 64, 128, 128],
        num_channels_skip=[4, 4, 4, 4, 4],
        filter_size_down=3, filter_size_up=3, filter_skip_size=1,
        need_sigmoid=True, need_tanh=False, need_bias=True,
        pad='zero', upsample_mode='bilinear', downsample_mode='stride', act_fun='LeakyReLU',
        need1x1_up=True):
    assert len(num_channels_down) == len(num_channels_up) == len(num_channels_skip)
    n = len(num_channels_down)
    up_modes, down_modes = _listify(upsample_mode, n), _listify(downsample_mode, n)
    k_down, k_up = _listify(filter_size_down, n), _listify(filter_size_up, n)

    def fill_level(level: nn.Sequential, i: int, cin: int) -> None:
        """Populate `level` (an empty Sequential) with scale i, recursing into the deeper scales."""
        last = i == n - 1
        c_skip, c_down, c_up = num_channels_skip[i], num_channels_down[i], num_channels_up[i]
        c_deep = c_down if last else num_channels_up[i + 1]

        deeper, branch = nn.Sequential(), nn.Sequential()
        level.add(Concat(1, branch, deeper) if c_skip != 0 else deeper)
        level.add(bn(c_skip + c_deep))

        if c_skip != 0:
            branch.add(conv(cin, c_skip, filter_skip_size, bias=need_bias, pad=pad))
            branch.add(bn(c_skip))
            branch.add(act(act_fun))

        deeper.add(conv(cin, c_down, k_down[i], 2, bias=need_bias, pad=pad, downsample_mode=down_modes[i]))
        deeper.add(bn(c_down))
        deeper.add(act(act_fun))
        deeper.add(conv(c_down, c_down, k_down[i], bias=need_bias, pad=pad))
        deeper.add(bn(c_down))
        deeper.add(act(act_fun))
        nxt = nn.Sequential()
        if not last:
            deeper.add(nxt)
        deeper.add(nn.Upsample(scale_factor=2, mode=up_modes[i]))

        level.add(conv(c_skip + c_deep, c_up, k_up[i], 1, bias=need_bias, pad=pad))
        level.add(bn(c_up))
        level.add(act(act_fun))
        if need1x1_up:
            level.add(conv(c_up, c_up, 1, bias=need_bias, pad=pad))
            level.add(bn(c_up))
            level.add(act(act_fun))
        if not last:
            fill_level(nxt, i + 1, c_down)

    # The default-argument network (the one the optimisation loop builds, models/networks.py:57) runs on the generator engine
    # tuned for it (csrc/generator.cu). Other configurations made of the same building blocks - inversion.py:21-25: 6 scales,
    # 32 input channels, 7x7 / 5x5 / 3x3 filters, reflection padding - run on the generalised engine (csrc/generator_x.cu).
    # Anything else (other activations / down-samplers / nearest up-sampling / no bias) is refused: evaluating it module by
    # module with torch / cuDNN would be a silent second backend (north_star: no multi-backend dispatch).
    is_default = (num_input_channels == 3 and num_output_channels == 3 and list(num_channels_down) == [16, 32, 64, 128, 128]
                  and list(num_channels_up) == [16, 32, 64, 128, 128] and list(num_channels_skip) == [4, 4, 4, 4, 4]
                  and k_down == [3] * 5 and k_up == [3] * 5 and filter_skip_size == 1 and need_sigmoid and need_bias
                  and pad == 'zero' and up_modes == ['bilinear'] * 5 and down_modes == ['stride'] * 5
                  and act_fun == 'LeakyReLU' and need1x1_up)
    if is_default:
        model = NativeSkip()
    else:
        filters = set(k_down) | set(k_up) | {filter_skip_size}
        widest = max([num_input_channels] + [num_channels_skip[i] + (num_channels_down[i] if i == n - 1 else num_channels_up[i + 1])
                                              for i in range(n)] + list(num_channels_down) + list(num_channels_up))
        supported = (act_fun == 'LeakyReLU' and need1x1_up and need_bias and not (need_tanh and not need_sigmoid)
                     and up_modes == ['bilinear'] * n and down_modes == ['stride'] * n and pad in ('zero', 'reflection')
                     and all(c > 0 for c in num_channels_skip) and filters <= {1, 3, 5, 7} and 1 <= n <= 8
                     and widest <= 160 and 1 <= num_output_channels <= 16)
        if not supported:
            raise NotImplementedError("splice_b200's generator engines implement skip() networks built from strided convs with bias "
                                      "(filter sizes 1/3/5/7, zero or reflection padding), BatchNorm, LeakyReLU, bilinear "
                                      "up-sampling, non-empty skip branches and 1x1 up convs, with at most 8 scales and 160 "
                                      "channels per tensor; use the reference's models/unet for other configurations")
        model = NativeSkipX(dict(num_input_channels=num_input_channels, num_output_channels=num_output_channels,
                                 num_channels_down=list(num_channels_down), num_channels_up=list(num_channels_up),
                                 num_channels_skip=list(num_channels_skip), filter_size_down=k_down, filter_size_up=k_up,
                                 filter_skip_size=filter_skip_size, pad=pad, need_sigmoid=bool(need_sigmoid)))
    fill_level(model, 0, num_input_channels)
    model.add(conv(num_channels_up[0], num_output_channels, 1, bias=need_bias, pad=pad))
    if need_sigmoid:
        model.add(nn.Sigmoid())
    elif need_tanh:
        model.add(nn.Tanh())
    return model
