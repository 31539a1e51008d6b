"""Drop-in for the reference's `models/extractor.py` (VitExtractor + attn_cosine_sim), backed by the
sm_100a ViT engine instead of a hooked torch module.

Same public names, argument orders and return shapes as /root/reference/models/extractor.py:4-163. The
reference registers 48 forward hooks per call and returns *lists of all 12 layers'* tensors; the engine
computes those taps directly (fp32 exports of the qkv GEMM / residual stream). The fused loss path
(util/losses.py) does not go through these list-returning methods.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Optional

import torch

from ..engine import DINO_ARCH, VitEngine


def attn_cosine_sim(x, eps=1e-08):
    """ref extractor.py:4-9 — x [1,1,t,D] -> [1,t,t]. Runs on the engine's Gram kernels when an engine is
    alive on x's device with matching D, otherwise raises (no CPU fallback)."""
    x = x[0]
    eng = VitExtractor._engine_for(x)
    if eng is None:
        raise RuntimeError("attn_cosine_sim needs a VitExtractor (engine) on a CUDA device with matching width")
    return eng.keys_self_sim(x[0].detach().float().contiguous())[None]


class _TapFn(torch.autograd.Function):
    """All-layer taps of one image as an autograd node (ref: inversion.py:33-39 back-propagates through
    get_feature_from_input / get_keys_from_input into the generator). Forward = one engine pass that keeps its activations
    (slot 6); backward = the engine's dgrad-only pass started from whatever per-layer tap gradients arrive."""

    @staticmethod
    def forward(ctx, input_img, extractor, kind):
        eng = extractor.engine
        img = input_img[0]
        res = eng.forward([img], (img.shape[1], img.shape[2]), n_grad=1, slot=_TapFn.SLOT, want_keys=False, want_cls=False,
                          want_all_qkv=(kind == "qkv"), want_all_blocks=(kind == "block"), pre_normalized=True)
        _TapFn.token += 1
        ctx.eng, ctx.kind, ctx.token = eng, kind, _TapFn.token
        ctx.set_materialize_grads(False)
        return tuple(res[kind][i] for i in range(12))

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.token != _TapFn.token:
            raise RuntimeError("VitExtractor: another differentiable tap call replaced this pass's activations before "
                               "backward() (one differentiable call may be alive at a time)")
        if all(g is None for g in gouts):
            return None, None, None
        kw = {"dblocks": list(gouts)} if ctx.kind == "block" else {"dqkvs": list(gouts)}
        g = ctx.eng.backward(_TapFn.SLOT, None, None, **kw)[0]
        return g[None], None, None


_TapFn.SLOT = 6
_TapFn.token = 0


class VitExtractor:
    BLOCK_KEY = 'block'
    ATTN_KEY = 'attn'
    PATCH_IMD_KEY = 'patch_imd'
    QKV_KEY = 'qkv'
    KEY_LIST = [BLOCK_KEY, ATTN_KEY, PATCH_IMD_KEY, QKV_KEY]

    # engines alive in this process, weakly held: attn_cosine_sim (a module-level function in the reference) needs one,
    # but a registry must not keep ~0.7 GB of packed weights + activation pools alive after its LossG is dropped
    _engines: "weakref.WeakSet[VitEngine]" = weakref.WeakSet()
    _engine_order = 0

    def __init__(self, model_name, device, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 packed: Optional[torch.Tensor] = None):
        """ref extractor.py:19-29. `state_dict` (optional, extension): DINO weights to use instead of
        `torch.hub.load('facebookresearch/dino:main', model_name)`; `packed`: the same weights as one flat
        fp32 buffer (engine.pack_vit_weights; what rank 0 broadcasts over NCCL); with SPLICE_B200_RANDOM_DINO=1 a seeded
        random DINO-style init is used when the hub is unreachable (benchmarks / offline tests)."""
        if model_name not in DINO_ARCH:
            raise NotImplementedError(f"model {model_name!r} is not a DINO ViT supported by splice_b200")
        self.model_name = model_name
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("splice_b200.VitExtractor runs on sm_100a only; there is no CPU fallback")
        self.model = None
        if state_dict is None and packed is None:
            if os.environ.get("SPLICE_B200_RANDOM_DINO", "0") == "1":
                from ..dino_init import random_dino_state_dict

                state_dict = random_dino_state_dict(model_name)
            else:
                # the hub module is only the source of the weights: packed on the host side, never kept on the GPU
                # (`.model` stays available on the CPU for code that inspects it, e.g. `.model.state_dict()`)
                self.model = torch.hub.load('facebookresearch/dino:main', model_name)
                self.model.eval()
                state_dict = self.model.state_dict()
        self.engine = VitEngine(model_name, state_dict, self.device, packed=packed)
        VitExtractor._engine_order += 1
        self.engine._registry_order = VitExtractor._engine_order
        VitExtractor._engines.add(self.engine)
        self.hook_handlers = []
        self.layers_dict = {key: list(range(12)) for key in VitExtractor.KEY_LIST}
        self.outputs_dict = {key: [] for key in VitExtractor.KEY_LIST}

    @classmethod
    def _engine_for(cls, x: torch.Tensor) -> Optional[VitEngine]:
        def idx(d):
            return d.index if d.index is not None else torch.cuda.current_device()

        best = None
        for e in list(cls._engines):
            if e.device.type == x.device.type == "cuda" and idx(e.device) == idx(x.device) and e.dim == x.shape[-1]:
                if best is None or e._registry_order > best._registry_order:
                    best = e
        return best

    # ---- taps (ref extractor.py:81-103) ------------------------------------------------------------
    def _run(self, input_img, **want):
        if input_img.dim() != 4 or input_img.shape[0] != 1:
            raise ValueError("VitExtractor expects a [1,3,h,w] batch (ref extractor.py:143 requires batch 1)")
        _, _, h, w = input_img.shape
        return self.engine.forward_normalized(input_img[0], **want)

    @staticmethod
    def _wants_grad(input_img) -> bool:
        return torch.is_grad_enabled() and input_img.requires_grad

    def _taps_with_grad(self, input_img, kind):
        if input_img.dim() != 4 or input_img.shape[0] != 1:
            raise ValueError("VitExtractor expects a [1,3,h,w] batch (ref extractor.py:143 requires batch 1)")
        if input_img.dtype != torch.float32 or not input_img.is_contiguous():
            input_img = input_img.float().contiguous()
        return list(_TapFn.apply(input_img, self, kind))

    def get_feature_from_input(self, input_img):  # List([B, N, D])
        if self._wants_grad(input_img):      # differentiable, like the reference's hooked module (inversion.py:35)
            return self._taps_with_grad(input_img, "block")
        res = self._run(input_img, want_all_blocks=True)
        return [res["block"][i] for i in range(12)]

    def get_qkv_feature_from_input(self, input_img):
        if self._wants_grad(input_img):
            return self._taps_with_grad(input_img, "qkv")
        res = self._run(input_img, want_all_qkv=True)
        return [res["qkv"][i] for i in range(12)]

    def get_attn_feature_from_input(self, input_img):
        """Materialising compatibility path: post-softmax probabilities [1,H,t,t] per layer (ref :97-103).
        Not used by the optimisation loop (the fused attention kernel never forms them)."""
        res = self._run(input_img, want_all_qkv=True)
        out = []
        H = self.get_head_num()
        for i in range(12):
            qkv = res["qkv"][i]
            t = qkv.shape[1]
            q, k, _ = qkv.reshape(1, t, 3, H, -1).permute(2, 0, 3, 1, 4)
            out.append(((q @ k.transpose(-2, -1)) * q.shape[-1] ** -0.5).softmax(dim=-1))
        return out

    # ---- shape helpers (ref extractor.py:105-130) ---------------------------------------------------
    def get_patch_size(self):
        return 8 if "8" in self.model_name else 16

    def get_width_patch_num(self, input_img_shape):
        b, c, h, w = input_img_shape
        return w // self.get_patch_size()

    def get_height_patch_num(self, input_img_shape):
        b, c, h, w = input_img_shape
        return h // self.get_patch_size()

    def get_patch_num(self, input_img_shape):
        return 1 + self.get_height_patch_num(input_img_shape) * self.get_width_patch_num(input_img_shape)

    def get_head_num(self):
        if "dino" in self.model_name:
            return 6 if "s" in self.model_name else 12
        return 6 if "small" in self.model_name else 12

    def get_embedding_dim(self):
        if "dino" in self.model_name:
            return 384 if "s" in self.model_name else 768
        return 384 if "small" in self.model_name else 768

    # ---- q/k/v slicing (ref extractor.py:132-151) ---------------------------------------------------
    def _split_qkv(self, qkv, input_img_shape, which):
        t = self.get_patch_num(input_img_shape)
        H = self.get_head_num()
        D = self.get_embedding_dim()
        return qkv.reshape(t, 3, H, D // H).permute(1, 2, 0, 3)[which]

    def get_queries_from_qkv(self, qkv, input_img_shape):
        return self._split_qkv(qkv, input_img_shape, 0)

    def get_keys_from_qkv(self, qkv, input_img_shape):
        return self._split_qkv(qkv, input_img_shape, 1)

    def get_values_from_qkv(self, qkv, input_img_shape):
        return self._split_qkv(qkv, input_img_shape, 2)

    def get_keys_from_input(self, input_img, layer_num):
        """[H,t,dh] keys of `layer_num` (ref :153-156). Layer 11 is served by the engine's direct fp32 export."""
        if layer_num in (11, -1) and not self._wants_grad(input_img):
            res = self._run(input_img)
            t = res["keys"].shape[1]
            H = self.get_head_num()
            return res["keys"][0].reshape(t, H, -1).permute(1, 0, 2)
        qkv_features = self.get_qkv_feature_from_input(input_img)[layer_num]
        return self.get_keys_from_qkv(qkv_features, input_img.shape)

    def get_keys_self_sim_from_input(self, input_img, layer_num):
        """[1,t,t] cosine self-similarity of the concatenated keys (ref :158-163)."""
        keys = self.get_keys_from_input(input_img, layer_num=layer_num)
        h, t, d = keys.shape
        concatenated_keys = keys.transpose(0, 1).reshape(t, h * d).contiguous()
        return self.engine.keys_self_sim(concatenated_keys)[None]
