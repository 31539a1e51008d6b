"""Random DINO-style initialisation of a ViT state dict (names/shapes of facebookresearch/dino checkpoints).

Used when no pretrained hub cache is reachable (there is no network in the build/bench environment): the
bench contract allows "random-init weights of that architecture"; timing and parity do not depend on the
weight values. Real runs load the hub checkpoint exactly like the reference (models/extractor.py:20)."""
from __future__ import annotations

from typing import Dict

import torch

from .engine import DEPTH, DINO_ARCH


def random_dino_state_dict(model_name: str, seed: int = 1234, device: str = "cpu") -> Dict[str, torch.Tensor]:
    patch, dim, _ = DINO_ARCH[model_name]
    g = torch.Generator(device="cpu").manual_seed(seed)

    def tn(*shape):
        t = torch.empty(*shape)
        torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
        return t

    n_pos = 1 + (224 // patch) ** 2
    fan_in = 3 * patch * patch
    bound = 1.0 / fan_in ** 0.5
    sd = {
        "cls_token": tn(1, 1, dim),
        "pos_embed": tn(1, n_pos, dim),
        "patch_embed.proj.weight": (torch.rand(dim, 3, patch, patch, generator=g) * 2 - 1) * bound,
        "patch_embed.proj.bias": (torch.rand(dim, generator=g) * 2 - 1) * bound,
    }
    for i in range(DEPTH):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = torch.ones(dim), torch.zeros(dim)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = tn(3 * dim, dim), torch.zeros(3 * dim)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = tn(dim, dim), torch.zeros(dim)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = torch.ones(dim), torch.zeros(dim)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = tn(4 * dim, dim), torch.zeros(4 * dim)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = tn(dim, 4 * dim), torch.zeros(dim)
    sd["norm.weight"], sd["norm.bias"] = torch.ones(dim), torch.zeros(dim)
    return {k: v.to(device) for k, v in sd.items()}
