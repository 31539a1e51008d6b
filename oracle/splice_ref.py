"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the Splice hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module. It restates, as pure torch-fp32 functions over plain state dicts, the algorithm of the reference's
per-image optimisation step; each function cites the reference file:line it follows (paths relative to the
reference repo root, /root/reference in the build container). `oracle/make_golden.py` validates every
function here against the unmodified reference imported from /root/reference and writes tests/golden/.

PARITY PINNING: the reference ships no tests/golden vectors. This restatement is pinned
by outputs of the reference itself run in the build container (tests/golden/*.pt, generator:
oracle/make_golden.py); the DINO dependency is pinned only by the reference's call sites (oracle/dino_vit.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # util/losses.py:19
IMAGENET_STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------------------------------
# preprocessing — util/losses.py:19-24 (Resize(dino_global_patch_size, max_size=480) -> Normalize)
# ------------------------------------------------------------------------------------------------
def resized_hw(h: int, w: int, size: int, max_size: int = 480) -> Tuple[int, int]:
    """Output (h, w) of torchvision `Resize(size:int, max_size)`: short side -> size (long side truncated),
    then clamped so that the long side does not exceed max_size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    if max_size is not None and new_long > max_size:
        new_short, new_long = int(max_size * new_short / new_long), max_size
    return (new_long, new_short) if w <= h else (new_short, new_long)


def aa_bilinear_matrix(n_in: int, n_out: int) -> Tensor:
    """Dense [n_out, n_in] weights of ATen's antialiased bilinear resampling along one axis
    (`_upsample_bilinear2d_aa`, align_corners=False) — the kernel torchvision>=0.17 `Resize` dispatches to for
    tensors. Triangle filter whose support widens by the scale when down-sampling;
    for up-sampling it degenerates to plain bilinear."""
    scale = n_in / n_out
    support = scale if scale >= 1.0 else 1.0
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    W = torch.zeros(n_out, n_in, dtype=torch.float64)
    for i in range(n_out):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), n_in)
        js = torch.arange(lo, hi, dtype=torch.float64)
        w = (1.0 - ((js - center + 0.5) * inv).abs()).clamp_min(0.0)
        W[i, lo:hi] = w / w.sum()
    return W.float()


def global_transform(img: Tensor, size: int = 224, max_size: int = 480) -> Tensor:
    """img [3,h,w] in [0,1] -> normalised [3,h',w'] (losses.py:19-24,77-78). Identity resize when the size
    already matches (torchvision returns the input unchanged)."""
    _, h, w = img.shape
    nh, nw = resized_hw(h, w, size, max_size)
    if (nh, nw) != (h, w):
        img = torch.einsum("ih,chw,jw->cij", aa_bilinear_matrix(h, nh).to(img), img, aa_bilinear_matrix(w, nw).to(img))
    mean = torch.tensor(IMAGENET_MEAN, dtype=img.dtype, device=img.device).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=img.dtype, device=img.device).view(3, 1, 1)
    return (img - mean) / std


# ------------------------------------------------------------------------------------------------
# DINO ViT forward with the reference's feature taps — models/extractor.py:40-103 without the hooks
# ------------------------------------------------------------------------------------------------
def vit_dims(sd: Dict[str, Tensor]) -> Tuple[int, int, int]:
    """(patch, D, heads) from a DINO state dict (heads: dh = 64 for every DINO ViT)."""
    w = sd["patch_embed.proj.weight"]
    return w.shape[-1], w.shape[0], w.shape[0] // 64


def vit_pos_embed(sd: Dict[str, Tensor], h: int, w: int) -> Tensor:
    from .dino_vit import interpolate_pos_embed

    patch = sd["patch_embed.proj.weight"].shape[-1]
    return interpolate_pos_embed(sd["pos_embed"], patch, (h // patch) * (w // patch), h, w)


def vit_taps(sd: Dict[str, Tensor], img: Tensor, eps: float = 1e-6) -> Dict[str, List[Tensor]]:
    """img [1,3,h,w] -> {'block': 12 x [1,t,D] (pre-final-norm block outputs, extractor.py:51-55,81-87),
    'qkv': 12 x [1,t,3D] (extractor.py:63-67,89-95), 'attn': 12 x [1,H,t,t] post-softmax (extractor.py:57-61)}"""
    patch, D, H = vit_dims(sd)
    _, _, h, w = img.shape
    x = F.conv2d(img, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd["cls_token"].expand(x.shape[0], -1, -1), x], dim=1) + vit_pos_embed(sd, h, w)
    t = x.shape[1]
    taps: Dict[str, List[Tensor]] = {"block": [], "qkv": [], "attn": []}
    for i in range(12):
        p = f"blocks.{i}."
        a = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        qkv = F.linear(a, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        taps["qkv"].append(qkv)
        q, k, v = qkv.reshape(1, t, 3, H, D // H).permute(2, 0, 3, 1, 4)
        prob = ((q @ k.transpose(-2, -1)) * (D // H) ** -0.5).softmax(dim=-1)
        taps["attn"].append(prob)
        o = (prob @ v).transpose(1, 2).reshape(1, t, D)
        x = x + F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        a = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
        hdn = F.gelu(F.linear(a, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = x + F.linear(hdn, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        taps["block"].append(x)
    return taps


def keys_from_qkv(qkv: Tensor, heads: int) -> Tensor:
    """[1,t,3D] -> keys [H,t,dh] (extractor.py:139-144; batch must be 1)."""
    t, d3 = qkv.shape[1], qkv.shape[2]
    return qkv.reshape(t, 3, heads, d3 // 3 // heads).permute(1, 2, 0, 3)[1]


def attn_cosine_sim(x: Tensor, eps: float = 1e-8) -> Tensor:
    """x [1,1,t,D] -> cosine self-similarity [1,t,t] (extractor.py:4-9)."""
    x = x[0]
    n = x.norm(dim=2, keepdim=True)
    return (x @ x.transpose(1, 2)) / (n @ n.transpose(1, 2)).clamp(min=eps)


def keys_self_sim(sd: Dict[str, Tensor], img: Tensor, layer: int = 11) -> Tensor:
    """extractor.py:158-163."""
    _, _, H = vit_dims(sd)
    keys = keys_from_qkv(vit_taps(sd, img)["qkv"][layer], H)
    h, t, d = keys.shape
    return attn_cosine_sim(keys.transpose(0, 1).reshape(t, h * d)[None, None])


# ------------------------------------------------------------------------------------------------
# losses — util/losses.py:34-105
# ------------------------------------------------------------------------------------------------
def active_lambdas(cfg: dict, step: int, state: Dict[str, float] | None = None) -> Dict[str, float]:
    """λ schedule of LossG.update_lambda_config (losses.py:34-44). `state` carries the sticky part
    (global_ssim / identity switch on at step == cls_warmup and stay on)."""
    lam = state if state is not None else {
        "lambda_global_cls": cfg["lambda_global_cls"], "lambda_global_ssim": 0, "lambda_entire_ssim": 0,
        "lambda_entire_cls": 0, "lambda_global_identity": 0}
    if step == cfg["cls_warmup"]:
        lam["lambda_global_ssim"] = cfg["lambda_global_ssim"]
        lam["lambda_global_identity"] = cfg["lambda_global_identity"]
    on = step % cfg["entire_A_every"] == 0
    lam["lambda_entire_ssim"] = cfg["lambda_entire_ssim"] if on else 0
    lam["lambda_entire_cls"] = cfg["lambda_entire_cls"] if on else 0
    return lam


def ssim_loss(sd, outputs: Tensor, inputs: Tensor, size: int = 224) -> Tensor:
    """calculate_global_ssim_loss (losses.py:74-83): per crop MSE of key self-similarities, summed."""
    loss = 0.0
    for a, b in zip(inputs, outputs):
        with torch.no_grad():
            target = keys_self_sim(sd, global_transform(a, size)[None])
        loss = loss + F.mse_loss(keys_self_sim(sd, global_transform(b, size)[None]), target)
    return loss


def cls_loss(sd, outputs: Tensor, inputs: Tensor, size: int = 224) -> Tensor:
    """calculate_crop_cls_loss (losses.py:85-94): MSE of block-11 [CLS] tokens, note zip(outputs, inputs)."""
    loss = 0.0
    for a, b in zip(outputs, inputs):
        cls = vit_taps(sd, global_transform(a, size)[None])["block"][-1][0, 0, :]
        with torch.no_grad():
            target = vit_taps(sd, global_transform(b, size)[None])["block"][-1][0, 0, :]
        loss = loss + F.mse_loss(cls, target)
    return loss


def id_loss(sd, outputs: Tensor, inputs: Tensor, size: int = 224) -> Tensor:
    """calculate_global_id_loss (losses.py:96-105): MSE of layer-11 keys [H,t,dh]."""
    _, _, H = vit_dims(sd)
    loss = 0.0
    for a, b in zip(inputs, outputs):
        with torch.no_grad():
            ka = keys_from_qkv(vit_taps(sd, global_transform(a, size)[None])["qkv"][11], H)
        kb = keys_from_qkv(vit_taps(sd, global_transform(b, size)[None])["qkv"][11], H)
        loss = loss + F.mse_loss(ka, kb)
    return loss


def loss_g(sd, cfg: dict, lam: Dict[str, float], outputs: Dict[str, Tensor], inputs: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """LossG.forward (losses.py:46-72) for already-updated lambdas."""
    size = cfg["dino_global_patch_size"]
    losses: Dict[str, Tensor] = {}
    total = 0
    if lam["lambda_global_ssim"] > 0:
        losses["loss_global_ssim"] = ssim_loss(sd, outputs["x_global"], inputs["A_global"], size)
        total = total + losses["loss_global_ssim"] * lam["lambda_global_ssim"]
    if lam["lambda_entire_ssim"] > 0:
        losses["loss_entire_ssim"] = ssim_loss(sd, outputs["x_entire"], inputs["A"], size)
        total = total + losses["loss_entire_ssim"] * lam["lambda_entire_ssim"]
    if lam["lambda_entire_cls"] > 0:
        losses["loss_entire_cls"] = cls_loss(sd, outputs["x_entire"], inputs["B_global"], size)
        total = total + losses["loss_entire_cls"] * lam["lambda_entire_cls"]
    if lam["lambda_global_cls"] > 0:
        losses["loss_global_cls"] = cls_loss(sd, outputs["x_global"], inputs["B_global"], size)
        total = total + losses["loss_global_cls"] * lam["lambda_global_cls"]
    if lam["lambda_global_identity"] > 0:
        losses["loss_global_id_B"] = id_loss(sd, outputs["y_global"], inputs["B_global"], size)
        total = total + losses["loss_global_id_B"] * lam["lambda_global_identity"]
    losses["loss"] = total
    return losses


# ------------------------------------------------------------------------------------------------
# generator — models/unet/skip.py:4-102 with default arguments, models/unet/common.py:11-42,76-124
# ------------------------------------------------------------------------------------------------
N_SCALES = 5
CH_DOWN = (16, 32, 64, 128, 128)
CH_UP = (16, 32, 64, 128, 128)
CH_SKIP = 4


def g_prefix(i: int) -> str:
    """state_dict prefix of scale i in the reference's nested nn.Sequential naming (common.py:5-8)."""
    return "1.1.7." * i


def _bn_train(x: Tensor, sd, key: str, eps: float = 1e-5) -> Tensor:
    """BatchNorm2d in training mode (common.py:95-96; .eval() is never called): batch statistics, biased var."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * sd[key + ".weight"].view(1, -1, 1, 1) + sd[key + ".bias"].view(1, -1, 1, 1)


def _conv(x: Tensor, sd, key: str, stride: int = 1) -> Tensor:
    w = sd[key + ".weight"]
    return F.conv2d(x, w, sd[key + ".bias"], stride=stride, padding=(w.shape[-1] - 1) // 2)  # common.py:113,120


def _lrelu(x: Tensor) -> Tensor:
    return F.leaky_relu(x, 0.2)  # common.py:82


def _center_crop_cat(a: Tensor, b: Tensor) -> Tensor:
    """Concat(dim=1) with centre-crop to the smaller spatial size (common.py:19-39)."""
    th, tw = min(a.shape[2], b.shape[2]), min(a.shape[3], b.shape[3])

    def crop(t):
        dh, dw = (t.shape[2] - th) // 2, (t.shape[3] - tw) // 2
        return t[:, :, dh:dh + th, dw:dw + tw]

    return torch.cat([crop(a), crop(b)], dim=1)


def _scale(x: Tensor, sd, i: int) -> Tensor:
    p = g_prefix(i)
    s = _lrelu(_bn_train(_conv(x, sd, p + "1.0.1.0"), sd, p + "1.0.2"))            # skip branch, skip.py:59-62
    d = _lrelu(_bn_train(_conv(x, sd, p + "1.1.1.0", stride=2), sd, p + "1.1.2"))   # skip.py:66-69
    d = _lrelu(_bn_train(_conv(d, sd, p + "1.1.4.0"), sd, p + "1.1.5"))             # skip.py:71-73
    if i < N_SCALES - 1:
        d = _scale(d, sd, i + 1)                                                    # skip.py:81
    d = F.interpolate(d, scale_factor=2, mode="bilinear")                           # skip.py:84
    c = _bn_train(_center_crop_cat(s, d), sd, p + "2")                              # skip.py:52-57
    c = _lrelu(_bn_train(_conv(c, sd, p + "3.0"), sd, p + "4"))                     # skip.py:86-88
    return _lrelu(_bn_train(_conv(c, sd, p + "6.0"), sd, p + "7"))                  # skip.py:90-93


def generator_forward(sd: Dict[str, Tensor], x: Tensor) -> Tensor:
    """netG(x): [n,3,h,w] -> [n,3,h',w'] in (0,1) (skip.py:97-99: final 1x1 conv + sigmoid)."""
    return torch.sigmoid(_conv(_scale(x, sd, 0), sd, "9.0"))


def generator_param_keys() -> List[str]:
    """Parameter names in `netG.parameters()` order (the order Adam sees them, util/util.py:30)."""
    keys: List[str] = []

    def rec(i):
        p = g_prefix(i)
        for k in ("1.0.1.0", "1.0.2", "1.1.1.0", "1.1.2", "1.1.4.0", "1.1.5"):
            keys.extend([p + k + ".weight", p + k + ".bias"])
        if i < N_SCALES - 1:
            rec(i + 1)
        for k in ("2", "3.0", "4", "6.0", "7"):
            keys.extend([p + k + ".weight", p + k + ".bias"])

    rec(0)
    keys.extend(["9.0.weight", "9.0.bias"])
    return keys


def generator_init_state(seed: int = 0, init_gain: float = 0.02) -> Dict[str, Tensor]:
    """A freshly initialised netG state (parameters only; BatchNorm running buffers never influence outputs because
    .eval() is never called) with the reference's init law - Xavier-normal(gain) conv weights, zero biases, BatchNorm
    weight ~ N(1, gain), bias 0 (networks.py:24-47) - drawn from a private generator. NOT the reference's RNG stream
    (that is pinned by tests/golden via the product's define_G); used where only the shapes/statistics matter
    (bench.py's reference arm must not import the product package)."""
    g = torch.Generator().manual_seed(seed)
    cin = (3,) + CH_DOWN[:-1]
    sd: Dict[str, Tensor] = {}

    def conv(key, co, ci, k):
        std = init_gain * math.sqrt(2.0 / (ci * k * k + co * k * k))
        sd[key + ".weight"] = torch.randn(co, ci, k, k, generator=g) * std
        sd[key + ".bias"] = torch.zeros(co)

    def bn(key, c):
        sd[key + ".weight"] = 1.0 + init_gain * torch.randn(c, generator=g)
        sd[key + ".bias"] = torch.zeros(c)

    for i in range(N_SCALES):
        p = g_prefix(i)
        deep = CH_DOWN[i] if i == N_SCALES - 1 else CH_UP[i + 1]
        conv(p + "1.0.1.0", CH_SKIP, cin[i], 1); bn(p + "1.0.2", CH_SKIP)
        conv(p + "1.1.1.0", CH_DOWN[i], cin[i], 3); bn(p + "1.1.2", CH_DOWN[i])
        conv(p + "1.1.4.0", CH_DOWN[i], CH_DOWN[i], 3); bn(p + "1.1.5", CH_DOWN[i])
        bn(p + "2", CH_SKIP + deep)
        conv(p + "3.0", CH_UP[i], CH_SKIP + deep, 3); bn(p + "4", CH_UP[i])
        conv(p + "6.0", CH_UP[i], CH_UP[i], 1); bn(p + "7", CH_UP[i])
    conv("9.0", 3, CH_UP[0], 1)
    return {k: sd[k] for k in generator_param_keys()}


# ------------------------------------------------------------------------------------------------
# Adam — util/util.py:28-32 -> torch.optim.Adam (lr 2e-3, betas (0, 0.99), eps 1e-8, no weight decay)
# ------------------------------------------------------------------------------------------------
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float, b2: float,
              eps: float = 1e-8) -> None:
    """In place; `step` is the 1-based step count after the increment (torch/optim/adam.py single-tensor path)."""
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
