"""ORACLE (test infrastructure, never on the product path) — stand-in for the absent third-party dependency.

The reference loads its feature extractor with `torch.hub.load('facebookresearch/dino:main', name)`
(/root/reference/models/extractor.py:20): a floating, un-pinned branch of facebookresearch/dino
(`hubconf.py` + `vision_transformer.py`) plus weights fetched from dl.fbaipublicfiles.com. Neither is
reachable offline, so this file restates the published architecture of DINO's `VisionTransformer` as
constrained by the reference's own call sites:

  * `model.blocks` is an iterable of 12 blocks                        (extractor.py:32,41)
  * `block.attn.qkv` is `Linear(D, 3D)`, output reshaped (t,3,H,dh)    (extractor.py:46,143)
  * `block.attn.attn_drop` receives the post-softmax probabilities    (extractor.py:44-45)
  * `block.attn(...)` returns a tuple whose [0] is the projected out  (extractor.py:48-49,77)
  * a block's output is the pre-final-norm residual stream [B,t,D]    (extractor.py:42-43, losses.py:90)

Weights are NOT the pretrained DINO weights (unobtainable here): `build()` seeds DINO's own init scheme
(trunc-normal sigma=0.02 for linears / cls / pos, LayerNorm (1,0), default Conv2d init for the patch embed).
Parity between splice_b200 and the reference is weight-agnostic — both sides are handed the same state_dict.

PARITY PINNING: the reference has no tests or golden vectors; this stand-in is pinned only by
the reference's call sites above. "parity unpinned" at the DINO boundary until a hub cache is available.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# name -> (patch, embed dim, heads); depth 12, mlp ratio 4, qkv bias, LayerNorm eps 1e-6 for all of them
ARCH = {
    "dino_vits16": (16, 384, 6),
    "dino_vits8": (8, 384, 6),
    "dino_vitb16": (16, 768, 12),
    "dino_vitb8": (8, 768, 12),
}
DEPTH = 12
LN_EPS = 1e-6
TRAIN_RES = 224  # pos_embed grid is (224/patch)^2


class _Attn(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        b, t, d = x.shape
        q, k, v = self.qkv(x).reshape(b, t, 3, self.num_heads, d // self.num_heads).permute(2, 0, 3, 1, 4)
        p = self.attn_drop(((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1))
        y = self.proj_drop(self.proj((p @ v).transpose(1, 2).reshape(b, t, d)))
        return y, p


class _Mlp(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(4 * dim, dim)
        self.drop = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _Mlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))[0]
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, patch: int, dim: int):
        super().__init__()
        self.patch_size = patch
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class DinoViT(nn.Module):
    def __init__(self, patch: int, dim: int, heads: int):
        super().__init__()
        self.embed_dim = dim
        self.patch_embed = _PatchEmbed(patch, dim)
        n = (TRAIN_RES // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, dim))
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(DEPTH)])
        self.norm = nn.LayerNorm(dim, eps=LN_EPS)
        self.head = nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def interpolate_pos_encoding(self, ntok: int, h: int, w: int):
        return interpolate_pos_embed(self.pos_embed, self.patch_embed.patch_size, ntok, h, w)

    def prepare_tokens(self, img):
        b, _, h, w = img.shape
        x = self.patch_embed(img)
        x = torch.cat([self.cls_token.expand(b, -1, -1), x], dim=1)
        return self.pos_drop(x + self.interpolate_pos_encoding(x.shape[1] - 1, h, w))

    def forward(self, img):
        x = self.prepare_tokens(img)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]


def interpolate_pos_embed(pos_embed: torch.Tensor, patch: int, ntok: int, h: int, w: int) -> torch.Tensor:
    """DINO's bicubic resampling of the learned patch position grid for inputs whose token grid differs from
    the training grid (or is not square). `scale_factor` carries DINO's +0.1 fudge against rounding."""
    n = pos_embed.shape[1] - 1
    if ntok == n and h == w:
        return pos_embed
    dim = pos_embed.shape[-1]
    side = int(math.sqrt(n))
    gh, gw = h // patch, w // patch
    grid = pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, scale_factor=((gh + 0.1) / side, (gw + 0.1) / side), mode="bicubic")
    assert grid.shape[-2] == gh and grid.shape[-1] == gw
    return torch.cat([pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, dim)], dim=1)


def build(name: str, seed: int = 1234, stats: str = "init") -> DinoViT:
    """Deterministic stand-in model; the global torch RNG state is saved and restored around construction
    so that seeding of the caller's own stream (train.py:29-31) is not disturbed.

    stats="init": DINO's initialisation (near-uniform softmax, no outlier channels: the benign case).
    stats="trained": the same model re-scaled to the activation statistics a *trained* DINO ViT shows, which is
    what stresses a reduced-precision implementation (see `apply_trained_statistics`)."""
    patch, dim, heads = ARCH[name]
    state = torch.get_rng_state()
    try:
        torch.manual_seed(seed)
        model = DinoViT(patch, dim, heads)
        if stats == "trained":
            apply_trained_statistics(model, seed + 1)
        elif stats != "init":
            raise ValueError(stats)
    finally:
        torch.set_rng_state(state)
    return model


@torch.no_grad()
def apply_trained_statistics(model: DinoViT, seed: int = 1235, logit_gain: float = 3.5, n_outlier: int = 3,
                             outlier_gain: float = 60.0) -> None:
    """Re-scales a freshly initialised stand-in so that its activations look like a trained DINO ViT's (the real
    weights cannot be fetched offline):
      * peaky attention: the query / key rows of every qkv projection are scaled by `logit_gain` each, so the
        pre-softmax logits have a standard deviation of a few units (low-entropy softmax rows; with the plain init
        they are ~0.3 and every row is near-uniform);
      * massive activations: `n_outlier` residual-stream channels receive a large constant through the fc2 bias of
        block 1 (+-`outlier_gain`, i.e. 50-100x the typical channel) and larger fc2 rows in the later blocks - the
        "few channels dominate the LayerNorm statistics" regime of trained ViTs; the LayerNorm gains of those
        channels are small, as in trained models;
      * every bias is non-zero and the LayerNorm gains are not 1.
    Both sides of a parity test receive the resulting state_dict, so this changes the *stress*, not the contract."""
    g = torch.Generator().manual_seed(seed)
    dim = model.embed_dim

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    outl = torch.randperm(dim, generator=g)[:n_outlier]
    sign = torch.where(torch.rand(n_outlier, generator=g) < 0.5, -1.0, 1.0)
    model.patch_embed.proj.bias.add_(rn(dim, std=0.05))
    for i, blk in enumerate(model.blocks):
        w = blk.attn.qkv.weight
        w[:2 * dim] *= logit_gain * (384.0 / dim) ** 0.5   # q and k rows; logit variance grows with dim at fixed init std
        blk.attn.qkv.bias.copy_(rn(3 * dim, std=0.1))
        blk.attn.proj.bias.copy_(rn(dim, std=0.05))
        blk.mlp.fc1.bias.copy_(rn(4 * dim, std=0.1))
        blk.mlp.fc2.bias.copy_(rn(dim, std=0.05))
        # once the massive channels exist (from block 1's output on) they dominate each row's variance; trained
        # models compensate with larger LayerNorm gains on the ordinary channels
        comp = (1.0 + n_outlier * outlier_gain ** 2 / dim) ** 0.5 if i >= 2 else 1.0
        for nrm in (blk.norm1, blk.norm2):
            nrm.weight.copy_(comp * (1.0 + rn(dim, std=0.25)).abs().clamp_min(0.05))
            nrm.bias.copy_(rn(dim, std=0.1))
            nrm.weight[outl] = 0.05 + 0.05 * torch.rand(n_outlier, generator=g)
        if i == 1:
            blk.mlp.fc2.bias[outl] = sign * outlier_gain
        if i >= 2:
            blk.mlp.fc2.weight[outl] *= 8.0
    model.norm.weight.copy_((1.0 + rn(dim, std=0.25)).abs().clamp_min(0.05))
    model.norm.bias.copy_(rn(dim, std=0.1))


def hub_load_standin(repo: str, name: str, *args, **kwargs) -> DinoViT:
    """Drop-in for `torch.hub.load('facebookresearch/dino:main', name)`."""
    assert "dino" in repo, repo
    return build(name)
