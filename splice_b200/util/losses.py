"""Drop-in for the reference's `util/losses.py` (LossG), computed by the sm_100a engine.

Public surface kept from /root/reference/util/losses.py:11-105: `LossG(cfg)` with `.extractor`,
`.global_transform`, `.lambdas`, `update_lambda_config(step)`, `forward(outputs, inputs) -> dict`, and the
three `calculate_*` methods. What changes is the execution plan of `forward`:

  reference                                           splice_b200
  ---------                                           -----------
  6 (10 on "entire" steps) batch-1 ViT forwards,      every distinct image of the step goes through ONE batched
  2 (4) of them duplicates (x_global and B_global     engine forward per ViT input size; generated images first so
  are pushed through the ViT twice)                   their activations are kept
  autograd graph through 3-5 ViT passes, incl.        the loss kernels emit d(keys)/d(cls); one dgrad-only engine
  weight gradients nobody reads                       backward per group returns d(total)/d(generated image)
  loss terms as separate torch ops                    fused loss kernels, device-side scalars

`forward` evaluates the objective AND its gradient w.r.t. the generated images eagerly; the returned
`losses['loss']` is attached to the autograd graph of `outputs[...]` through `_Objective`, so
`losses['loss'].backward()` (train.py:78) propagates into netG exactly as in the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from ..engine import VitEngine, weighted_total
from ..models.extractor import VitExtractor

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

SSIM, CLS, KEYS = 0, 1, 2
# order of the loss terms in the device-side `terms` vector and in LossG.forward (ref losses.py:51-69)
TERM_ORDER = ("loss_global_ssim", "loss_entire_ssim", "loss_entire_cls", "loss_global_cls", "loss_global_id_B")


class GlobalTransform:
    """Callable stand-in for `transforms.Compose([Resize(size, max_size=480), Normalize(ImageNet)])`
    (ref losses.py:19-24) for code that applies `criterion.global_transform(img)` itself (inversion.py,
    keys_self_sim_pca.py). Inside LossG.forward the same arithmetic is fused into the patchify kernel."""

    def __init__(self, engine: VitEngine, size: int, max_size: int = 480):
        self.engine, self.size, self.max_size = engine, size, max_size

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        from .. import _lib
        from .._lib import check, cur_stream, ptr

        batched = img.dim() == 4
        if batched and img.shape[0] != 1:
            raise ValueError("global_transform takes one image ([3,h,w] or [1,3,h,w])")
        x = (img[0] if batched else img).detach().float().contiguous()
        _, h, w = x.shape
        oh, ow = self.engine.vit_input_size(h, w, self.size, self.max_size)
        out = torch.empty(3, oh, ow, device=x.device)
        check(_lib.splice_resize_normalize(ptr(x), h, w, oh, ow, ptr(out), 1, cur_stream()), "splice_resize_normalize")
        return out[None] if batched else out


class _Objective(torch.autograd.Function):
    """total (device scalar, already computed) as a function of the generated images; the backward returns the
    gradients the engine produced in the same pass, scaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, total: torch.Tensor, grads: Tuple[Optional[torch.Tensor], ...], *gens: torch.Tensor):
        ctx.grads = grads
        return total.clone()

    @staticmethod
    def backward(ctx, gout):
        outs = []
        for g in ctx.grads:
            outs.append(None if g is None else g * gout)
        return (None, None, *outs)


class LossG(torch.nn.Module):

    def __init__(self, cfg, state_dict=None, packed=None):
        super().__init__()
        self.cfg = cfg
        self.extractor = VitExtractor(model_name=cfg['dino_model_name'], device=device, state_dict=state_dict,
                                      packed=packed)
        self.engine: VitEngine = self.extractor.engine
        self.global_transform = GlobalTransform(self.engine, cfg['dino_global_patch_size'])
        self.overlap_targets = True     # targets' ViT pass on a side stream, in the shadow of the generator forward
        self.keys_only_stop = True      # sequences only read through their layer-11 keys stop after the last qkv projection
        self.run_ahead_max_pixels = 80_000   # generated images up to this size: the targets' pass may run ahead of the main stream
        self._side = None
        self._targets_consumed = None   # event: the loss kernels of the last forward() have read the targets' features
        self.lambdas = dict(
            lambda_global_cls=cfg['lambda_global_cls'],
            lambda_global_ssim=0,
            lambda_entire_ssim=0,
            lambda_entire_cls=0,
            lambda_global_identity=0
        )

    def _side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    # ref losses.py:34-44
    def update_lambda_config(self, step):
        step = float(step)  # accepts the reference's 1-element (device) tensor; one host sync like the reference
        if step == self.cfg['cls_warmup']:
            self.lambdas['lambda_global_ssim'] = self.cfg['lambda_global_ssim']
            self.lambdas['lambda_global_identity'] = self.cfg['lambda_global_identity']
        entire = step % self.cfg['entire_A_every'] == 0
        self.lambdas['lambda_entire_ssim'] = self.cfg['lambda_entire_ssim'] if entire else 0
        self.lambdas['lambda_entire_cls'] = self.cfg['lambda_entire_cls'] if entire else 0

    # ---- the fused objective -----------------------------------------------------------------------
    def _plan(self, outputs, inputs):
        """[(term name, kind, lambda, generated batch, target batch)] for the active terms (ref losses.py:51-69)."""
        lam = self.lambdas
        plan = []
        if lam['lambda_global_ssim'] > 0:
            plan.append(("loss_global_ssim", SSIM, lam['lambda_global_ssim'], outputs['x_global'], inputs['A_global']))
        if lam['lambda_entire_ssim'] > 0:
            plan.append(("loss_entire_ssim", SSIM, lam['lambda_entire_ssim'], outputs['x_entire'], inputs['A']))
        if lam['lambda_entire_cls'] > 0:
            plan.append(("loss_entire_cls", CLS, lam['lambda_entire_cls'], outputs['x_entire'], inputs['B_global']))
        if lam['lambda_global_cls'] > 0:
            plan.append(("loss_global_cls", CLS, lam['lambda_global_cls'], outputs['x_global'], inputs['B_global']))
        if lam['lambda_global_identity'] > 0:
            plan.append(("loss_global_id_B", KEYS, lam['lambda_global_identity'], outputs['y_global'], inputs['B_global']))
        return plan

    def forward(self, outputs, inputs):
        self.update_lambda_config(inputs['step'])
        plan = self._plan(outputs, inputs)
        eng = self.engine
        size = self.cfg['dino_global_patch_size']

        # 1. distinct images of the step: (batch tensor id, crop index) -> sequence record
        seqs: Dict[Tuple[int, int], dict] = {}
        gen_batches: List[torch.Tensor] = []

        def seq_of(batch: torch.Tensor, i: int, is_gen: bool) -> dict:
            key = (id(batch), i)
            if key not in seqs:
                img = batch[i]
                oh, ow = eng.vit_input_size(img.shape[1], img.shape[2], size)
                seqs[key] = {"batch": batch, "i": i, "img": img, "hw": (oh, ow), "gen": False}
            if is_gen and batch.requires_grad and torch.is_grad_enabled():
                seqs[key]["gen"] = True
                if all(b is not batch for b in gen_batches):
                    gen_batches.append(batch)
            return seqs[key]

        pairs = []  # (term index, kind, lambda, gen seq, target seq)
        for name, kind, lam, gen, tgt in plan:
            for i in range(min(len(gen), len(tgt))):  # zip() semantics of the reference loops
                pairs.append((TERM_ORDER.index(name), kind, lam, seq_of(gen, i, True), seq_of(tgt, i, False)))
        # sequences read only through their layer-11 keys (ssim / identity terms: ref losses.py:74-83,96-105) stop after the
        # last layer's qkv projection; those a [CLS] term reads (losses.py:85-94) run the full depth and go first in a pass
        for _, kind, _, g, tg in pairs:
            if kind == CLS:
                g["full"] = tg["full"] = True

        # 2. one batched engine forward per ViT input size, generated (grad) sequences first
        groups: Dict[Tuple[int, int], List[dict]] = {}
        for s in seqs.values():
            groups.setdefault(s["hw"], []).append(s)
        if len(groups) > 3:
            raise NotImplementedError("more than 3 distinct ViT input sizes in one step")
        main = torch.cuda.current_stream()
        for g_idx, (hw, members) in enumerate(groups.items()):
            members.sort(key=lambda s: (not s["gen"], not s.get("full", False)))
            n_grad = sum(1 for s in members if s["gen"])
            n_full_of = (lambda part: sum(1 for s in part if s.get("full", False))) if self.keys_only_stop else (lambda part: None)
            slot = 2 * g_idx
            if self.overlap_targets and 0 < n_grad < len(members):
                # The targets (A_global, B_global, A) do not depend on netG: their no-grad pass goes to a side stream
                # that only waits for the inputs, so it runs in the shadow of the generator forward the main stream is
                # still busy with; the generated images follow on the main stream, in their own (kept) slot.
                gens, tgts = members[:n_grad], members[n_grad:]
                side = self._side_stream()
                ready = [getattr(s["batch"], "_splice_ready", None) for s in tgts]
                # Run-ahead only pays while the generator leaves SMs idle (224 px: its kernels are latency-bound). Once its
                # conv kernels are throughput-bound, a persistent GEMM CTA (~200 KB of shared memory) parked on an SM
                # leaves room for ONE conv CTA instead of several: measured 117 -> 65 it/s at 448 px and 48 -> 25 it/s on
                # the 1200x900 pair with run-ahead on. Above the threshold the pass is ordered after the main stream.
                big = max(s["img"].shape[1] * s["img"].shape[2] for s in gens) > self.run_ahead_max_pixels
                if big or any(e is None for e in ready):
                    side.wait_stream(main)
                else:
                    for e in {id(e): e for e in ready}.values():
                        side.wait_event(e)
                    # ... and for the previous step's loss kernels, which read the targets' features this pass overwrites.
                    # Nothing else ties it to the main stream: it may run under the previous step's backward passes.
                    if self._targets_consumed is not None:
                        side.wait_event(self._targets_consumed)
                for s in tgts:
                    s["batch"].record_stream(side)
                ft = eng.forward([s["img"] for s in tgts], hw, n_grad=0, slot=slot + 1, use_graph=True, stream=side.cuda_stream,
                                 n_full=n_full_of(tgts))
                fg = eng.forward([s["img"] for s in gens], hw, n_grad=n_grad, slot=slot, use_graph=True, n_full=n_full_of(gens))
                main.wait_stream(side)
                t = fg["keys"].shape[1]
                parts = [(gens, fg), (tgts, ft)]
            else:
                # one pass: full-depth sequences must be a PREFIX, so every sequence up to the last one a [CLS] term reads runs full
                last_full = max([j for j, s in enumerate(members) if s.get("full", False)], default=-1)
                feats = eng.forward([s["img"] for s in members], hw, n_grad=n_grad, slot=slot, use_graph=True,
                                    n_full=(last_full + 1) if self.keys_only_stop else None)
                t = feats["keys"].shape[1]
                parts = [(members, feats)]
            dkeys, dcls = eng.grad_buffers(slot, n_grad, t) if n_grad else (None, None)
            for part, feats in parts:
                for j, s in enumerate(part):
                    s.update(keys=feats["keys"][j], cls=feats["cls"][j], slot=slot, idx=j,
                             dkeys=dkeys[j] if s["gen"] else None, dcls=dcls[j] if s["gen"] else None)
            members[0]["group"] = {"slot": slot, "n_grad": n_grad, "dkeys": dkeys, "dcls": dcls, "members": members}

        # 3. loss kernels: value into terms[k] (summed over crops), gradient into the generated sequence's slot
        terms = torch.zeros(8, device=device)  # fresh per call: the returned per-term scalars are views of it
        total = torch.empty(1, device=device)
        scratch = torch.zeros(1, device=device)
        per_term_count: Dict[int, int] = {}
        for k, kind, lam, g, tg in pairs:
            out = terms[k:k + 1] if per_term_count.get(k, 0) == 0 else scratch
            if kind == SSIM:
                eng.loss_ssim(g["keys"], tg["keys"], float(lam), out, g["dkeys"])
            elif kind == CLS:
                eng.loss_mse(g["cls"], tg["cls"], float(lam), out, g["dcls"])
            else:
                eng.loss_mse(g["keys"], tg["keys"], float(lam), out, g["dkeys"])
            if out is scratch:
                terms[k:k + 1] += scratch
            per_term_count[k] = per_term_count.get(k, 0) + 1
        weights = [0.0] * len(TERM_ORDER)
        for name, kind, lam, _, _ in plan:
            weights[TERM_ORDER.index(name)] = float(lam)
        weighted_total(terms, weights, total)
        self._targets_consumed = torch.cuda.Event()
        self._targets_consumed.record(main)

        # 4. dgrad-only backward per group -> d(total)/d(generated image), assembled per generated batch
        grads_by_batch: Dict[int, torch.Tensor] = {}
        for hw, members in groups.items():
            grp = members[0]["group"]
            if grp["n_grad"] == 0:
                continue
            img_grads = eng.backward(grp["slot"], grp["dkeys"], grp["dcls"], use_graph=True)
            for s, gimg in zip(members[:grp["n_grad"]], img_grads):
                b = s["batch"]
                if id(b) not in grads_by_batch:
                    grads_by_batch[id(b)] = torch.zeros_like(b)
                grads_by_batch[id(b)][s["i"]] = gimg

        losses = {}
        for name, *_ in plan:
            losses[name] = terms[TERM_ORDER.index(name)]
        if gen_batches:
            grads = tuple(grads_by_batch.get(id(b)) for b in gen_batches)
            losses['loss'] = _Objective.apply(total[0], grads, *gen_batches)
        else:
            losses['loss'] = total[0]
        return losses

    # ---- per-term API of the reference (ref losses.py:74-105); same engine, one term at a time ------------
    def _single(self, kind, gen, tgt):
        plan = [("loss_global_ssim" if kind == SSIM else "loss_global_cls" if kind == CLS else "loss_global_id_B",
                 kind, 1.0, gen, tgt)]
        saved = self._plan
        try:
            self._plan = lambda outputs, inputs: plan  # type: ignore[assignment]
            saved_update = self.update_lambda_config
            self.update_lambda_config = lambda step: None  # type: ignore[assignment]
            return LossG.forward(self, {}, {"step": 0})['loss']
        finally:
            self._plan = saved  # type: ignore[assignment]
            self.update_lambda_config = saved_update  # type: ignore[assignment]

    def calculate_global_ssim_loss(self, outputs, inputs):
        return self._single(SSIM, outputs, inputs)

    def calculate_crop_cls_loss(self, outputs, inputs):
        return self._single(CLS, outputs, inputs)

    def calculate_global_id_loss(self, outputs, inputs):
        return self._single(KEYS, outputs, inputs)
