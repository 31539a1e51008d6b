"""Host-side handle of the C++/CUDA ViT engine (include/splice_b200.h: splice_vit_*, splice_loss_*).

PyTorch is used for device memory (tensors are allocation + lifetime) and streams only; every arithmetic
step of the hot path runs in libsplice_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, cur_stream, ptr

# model name -> (patch, D, heads); the reference derives these from the *name string*
# (models/extractor.py:105-130), every DINO ViT has depth 12 and head dim 64.
DINO_ARCH = {
    "dino_vits16": (16, 384, 6),
    "dino_vits8": (8, 384, 6),
    "dino_vitb16": (16, 768, 12),
    "dino_vitb8": (8, 768, 12),
}
DEPTH = 12
LN_EPS = 1e-6


def packed_key_order(depth: int = DEPTH) -> List[str]:
    """state_dict keys in the order of the packed weight buffer (include/splice_b200.h, SpliceVitDesc)."""
    keys = ["cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    for i in range(depth):
        p = f"blocks.{i}."
        keys += [p + "norm1.weight", p + "norm1.bias", p + "attn.qkv.weight", p + "attn.qkv.bias",
                 p + "attn.proj.weight", p + "attn.proj.bias", p + "norm2.weight", p + "norm2.bias",
                 p + "mlp.fc1.weight", p + "mlp.fc1.bias", p + "mlp.fc2.weight", p + "mlp.fc2.bias"]
    keys += ["norm.weight", "norm.bias"]
    return keys


def pack_vit_weights(state_dict: Dict[str, torch.Tensor], device: torch.device | str = "cuda") -> torch.Tensor:
    """One flat fp32 buffer holding the 150 DINO tensors — also the buffer rank 0 broadcasts over NCCL."""
    parts = [state_dict[k].detach().to(torch.float32).reshape(-1) for k in packed_key_order()]
    return torch.cat(parts).to(device).contiguous()


def interpolate_pos_embed(pos_embed: torch.Tensor, patch: int, h: int, w: int) -> torch.Tensor:
    """DINO's bicubic resampling of the position grid for a non-native token grid (one-off per input shape;
    weight preprocessing, not part of the per-step hot path). Restated from oracle/dino_vit.py's description
    of facebookresearch/dino `interpolate_pos_encoding`."""
    n = pos_embed.shape[1] - 1
    gh, gw = h // patch, w // patch
    if gh * gw == n and h == w:
        return pos_embed[0]
    side = int(math.sqrt(n))
    dim = pos_embed.shape[-1]
    grid = pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2)
    grid = torch.nn.functional.interpolate(grid, scale_factor=((gh + 0.1) / side, (gw + 0.1) / side), mode="bicubic")
    assert grid.shape[-2] == gh and grid.shape[-1] == gw
    return torch.cat([pos_embed[0, :1], grid.permute(0, 2, 3, 1).reshape(-1, dim)], dim=0).contiguous()


class VitEngine:
    """Owns one `splice_vit_*` context on the current CUDA device."""

    def __init__(self, model_name: str, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 device: torch.device | str = "cuda", packed: Optional[torch.Tensor] = None, gemm_impl: int = 0):
        """Either `state_dict` (the 150 DINO tensors) or `packed` (the flat fp32 buffer of pack_vit_weights, e.g.
        as received from the rank-0 NCCL broadcast) must be given."""
        if model_name not in DINO_ARCH:
            raise NotImplementedError(f"unsupported DINO model {model_name!r}")
        if not torch.cuda.is_available():
            raise RuntimeError("splice_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.model_name = model_name
        self.patch, self.dim, self.heads = DINO_ARCH[model_name]
        self.device = torch.device(device)
        self.gemm_impl = gemm_impl
        if packed is None:
            if state_dict is None:
                raise ValueError("VitEngine needs a state_dict or a packed weight buffer")
            packed = pack_vit_weights(state_dict, self.device)
        self.packed = packed.to(self.device, torch.float32).contiguous()
        self.n_pos = 1 + (224 // self.patch) ** 2          # DINO checkpoints are trained at 224 px
        self.pos_embed = self.packed[self.dim:self.dim + self.n_pos * self.dim].view(1, self.n_pos, self.dim)
        self.desc = _lib.SpliceVitDesc(self.patch, self.dim, self.heads, DEPTH, self.n_pos, LN_EPS)
        expect = _lib.splice_vit_packed_floats(C.byref(self.desc))
        if self.packed.numel() != expect:
            raise ValueError(f"packed ViT weights have {self.packed.numel()} floats, the engine expects {expect}")
        self._ctx = C.c_void_p()
        check(_lib.splice_vit_create(C.byref(self._ctx), C.byref(self.desc), ptr(self.packed), self.packed.numel(),
                                     cur_stream()), "splice_vit_create")
        self._pos_cache: Dict[Tuple[int, int], torch.Tensor] = {}
        self._slot_meta: Dict[int, dict] = {}
        self._out_cache: Dict[tuple, Dict[str, torch.Tensor]] = {}
        self._grad_cache: Dict[tuple, tuple] = {}

    def __del__(self):
        ctx = getattr(self, "_ctx", None)
        if ctx:
            try:
                torch.cuda.synchronize()
                _lib.splice_vit_destroy(ctx)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._ctx = None

    # ---- shapes ------------------------------------------------------------------------------------
    def vit_input_size(self, h: int, w: int, size: int, max_size: int = 480) -> Tuple[int, int]:
        oh, ow = C.c_int(), C.c_int()
        _lib.splice_resized_hw(h, w, size, max_size, C.byref(oh), C.byref(ow))
        return oh.value, ow.value

    def tokens(self, oh: int, ow: int) -> int:
        return 1 + (oh // self.patch) * (ow // self.patch)

    def _pos_for(self, oh: int, ow: int) -> Optional[torch.Tensor]:
        gh, gw = oh // self.patch, ow // self.patch
        if 1 + gh * gw == self.n_pos and gh == gw:
            return None
        key = (oh, ow)
        if key not in self._pos_cache:
            self._pos_cache[key] = interpolate_pos_embed(self.pos_embed, self.patch, oh, ow)
        return self._pos_cache[key]

    # ---- forward / backward ------------------------------------------------------------------------
    def forward(self, images: Sequence[torch.Tensor], out_hw: Tuple[int, int], n_grad: int = 0, slot: int = 0,
                want_keys: bool = True, want_cls: bool = True, want_all_qkv: bool = False,
                want_all_blocks: bool = False, pre_normalized: bool = False, use_graph: bool = False,
                stream: Optional[int] = None, n_full: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """images: fp32 CUDA tensors [3,h,w] in [0,1] (sizes may differ); all are resized to out_hw.
        `stream` (raw cudaStream_t, default: torch's current stream): passes in different slots may be in flight on
        different streams; the caller orders the streams (outputs are cached per slot, so no allocation happens here
        after the first call). `n_full`: only the first n_full images need the last block's output ('cls'); the others
        are keys-only and stop after the last layer's qkv projection (None = all images run the full depth).
        Returns {'keys': [S,t,D], 'cls': [S,D], 'qkv': [12,S,t,3D], 'block': [12,S,t,D]} (only the requested)."""
        S = len(images)
        oh, ow = out_hw
        t = self.tokens(oh, ow)
        arr = (_lib.SpliceImage * S)()
        keep = []
        for i, im in enumerate(images):
            if im.dim() != 3 or im.shape[0] != 3:
                raise ValueError("images must be [3,h,w]")
            im = im.detach()
            if im.dtype != torch.float32 or not im.is_contiguous():
                im = im.to(torch.float32).contiguous()
            keep.append(im)
            arr[i] = _lib.SpliceImage(im.data_ptr(), im.shape[1], im.shape[2])
        dev, D = self.device, self.dim
        ckey = (slot, S, t, want_keys, want_cls, want_all_qkv, want_all_blocks)
        out: Optional[Dict[str, torch.Tensor]] = self._out_cache.get(ckey) if use_graph else None
        if out is None:
            out = {}
            if want_keys:
                out["keys"] = torch.empty(S, t, D, device=dev)
            if want_cls:
                out["cls"] = torch.empty(S, D, device=dev)
            if want_all_qkv:
                out["qkv"] = torch.empty(DEPTH, S, t, 3 * D, device=dev)
            if want_all_blocks:
                out["block"] = torch.empty(DEPTH, S, t, D, device=dev)
            if use_graph:   # stable output addresses let the engine replay a captured CUDA graph (results are
                self._out_cache[ckey] = out   # overwritten by the next call with the same signature)
        pos = self._pos_for(oh, ow)
        a = _lib.SpliceVitForwardArgs()
        a.images, a.n_images, a.out_h, a.out_w = arr, S, oh, ow
        a.pos = ptr(pos)
        a.n_grad, a.slot = n_grad, slot
        a.keys32, a.cls32 = ptr(out.get("keys")), ptr(out.get("cls"))
        a.qkv32_all, a.block32_all = ptr(out.get("qkv")), ptr(out.get("block"))
        a.gemm_impl = self.gemm_impl
        a.pre_normalized = 1 if pre_normalized else 0
        a.use_graph = 1 if use_graph else 0
        a.n_full = 0 if n_full is None or n_full >= S else (-1 if n_full <= 0 else n_full)
        check(_lib.splice_vit_forward(self._ctx, C.byref(a), cur_stream() if stream is None else stream), "splice_vit_forward")
        self._slot_meta[slot] = {"shapes": [(im.shape[1], im.shape[2]) for im in keep[:n_grad]], "t": t, "keep": keep}
        return out

    def forward_normalized(self, img: torch.Tensor, **want) -> Dict[str, torch.Tensor]:
        """One already-normalised [3,h,w] image at ViT resolution (the VitExtractor API contract)."""
        return self.forward([img], (img.shape[1], img.shape[2]), n_grad=0, slot=7, pre_normalized=True, **want)

    def grad_buffers(self, slot: int, n_grad: int, t: int):
        """Zeroed (dkeys [n_grad,t,D], dcls [n_grad,D]) with stable addresses (graph replay in backward)."""
        key = (slot, n_grad, t)
        if key not in self._grad_cache:
            self._grad_cache[key] = (torch.zeros(n_grad, t, self.dim, device=self.device),
                                     torch.zeros(n_grad, self.dim, device=self.device))
        dk, dc = self._grad_cache[key]
        dk.zero_()
        dc.zero_()
        return dk, dc

    def backward(self, slot: int, dkeys: Optional[torch.Tensor], dcls: Optional[torch.Tensor],
                 use_graph: bool = False, dblocks: Optional[Sequence[Optional[torch.Tensor]]] = None,
                 dqkvs: Optional[Sequence[Optional[torch.Tensor]]] = None) -> List[torch.Tensor]:
        """d(loss)/d(image) for the first n_grad images of the forward held in `slot`. `dblocks` / `dqkvs`: per-layer
        gradients w.r.t. the all-layer taps ([n_grad,t,D] / [n_grad,t,3D] or None), for differentiable VitExtractor calls."""
        meta = self._slot_meta[slot]
        n = len(meta["shapes"])
        grads = [torch.empty(3, h, w, device=self.device) for (h, w) in meta["shapes"]]
        arr = (_lib.SpliceImage * n)()
        for i, g in enumerate(grads):
            arr[i] = _lib.SpliceImage(g.data_ptr(), g.shape[1], g.shape[2])
        a = _lib.SpliceVitBackwardArgs()
        a.slot = slot
        if dkeys is not None:
            assert dkeys.dtype == torch.float32 and dkeys.is_contiguous() and dkeys.numel() == n * meta["t"] * self.dim
        if dcls is not None:
            assert dcls.dtype == torch.float32 and dcls.is_contiguous() and dcls.numel() == n * self.dim
        a.dkeys32, a.dcls32, a.grads, a.gemm_impl = ptr(dkeys), ptr(dcls), arr, self.gemm_impl
        a.use_graph = 1 if use_graph else 0
        hold = []
        for name, lst, width in (("dblock32_layers", dblocks, self.dim), ("dqkv32_layers", dqkvs, 3 * self.dim)):
            if lst is None or all(g is None for g in lst):
                continue
            if len(lst) != DEPTH:
                raise ValueError(f"{name}: one entry per layer ({DEPTH}) expected")
            arr2 = (C.c_void_p * DEPTH)()
            for i, g in enumerate(lst):
                if g is None:
                    continue
                g = g.detach().to(torch.float32).contiguous()
                if g.numel() != n * meta["t"] * width:
                    raise ValueError(f"{name}[{i}] has {g.numel()} elements, expected {n * meta['t'] * width}")
                hold.append(g)
                arr2[i] = g.data_ptr()
            setattr(a, name, arr2)
            hold.append(arr2)
        check(_lib.splice_vit_backward(self._ctx, C.byref(a), cur_stream()), "splice_vit_backward")
        return grads

    # ---- profiling -------------------------------------------------------------------------------------
    PROFILE_CLASSES = ("gemm_tcgen05", "attention_fwd", "attention_bwd", "rowwise_layernorm", "preprocess")

    def profile_enable(self, on: bool) -> None:
        check(_lib.splice_vit_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self) -> Dict[str, dict]:
        n = len(self.PROFILE_CLASSES)
        arr = (_lib.SpliceProfileEntry * n)()
        check(_lib.splice_vit_profile_read(self._ctx, arr, n))
        return {name: {"count": arr[i].count, "ms": arr[i].ms, "flops": arr[i].flops, "bytes": arr[i].bytes}
                for i, name in enumerate(self.PROFILE_CLASSES)}

    # ---- losses ------------------------------------------------------------------------------------
    def loss_ssim(self, keys_x: torch.Tensor, keys_a: torch.Tensor, coef: float, loss_out: torch.Tensor,
                  dkeys_out: Optional[torch.Tensor]) -> None:
        t = keys_x.shape[-2]
        check(_lib.splice_loss_ssim(self._ctx, ptr(keys_x), ptr(keys_a), t, coef, ptr(dkeys_out), ptr(loss_out),
                                    self.gemm_impl, cur_stream()), "splice_loss_ssim")

    def loss_mse(self, a: torch.Tensor, b: torch.Tensor, coef: float, loss_out: torch.Tensor,
                 grad_out: Optional[torch.Tensor]) -> None:
        rows, cols = (a.shape[-2], a.shape[-1]) if a.dim() >= 2 else (1, a.shape[-1])
        check(_lib.splice_loss_mse(self._ctx, ptr(a), ptr(b), rows, cols, coef, ptr(grad_out), ptr(loss_out),
                                   cur_stream()), "splice_loss_mse")

    def keys_self_sim(self, keys: torch.Tensor) -> torch.Tensor:
        t = keys.shape[-2]
        out = torch.empty(t, t, device=self.device)
        check(_lib.splice_keys_self_sim(self._ctx, ptr(keys), t, ptr(out), self.gemm_impl, cur_stream()),
              "splice_keys_self_sim")
        return out


def weighted_total(terms: torch.Tensor, weights: Sequence[float], total_out: torch.Tensor) -> None:
    n = len(weights)
    w = (C.c_float * n)(*weights)
    check(_lib.splice_weighted_total(ptr(terms), w, n, ptr(total_out), cur_stream()), "splice_weighted_total")
