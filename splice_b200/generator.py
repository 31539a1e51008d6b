"""Host-side glue between the generator's nn.Module tree (parameters, BatchNorm buffers, state_dict) and the
native sm_100a generator engine (include/splice_b200.h: splice_gen_*).

The module tree built by models/unet/skip.py only OWNS the tensors; `NativeSkip.forward` sends them by pointer
to the engine. Parameter gradients are accumulated by the engine straight into the `.grad` buffers (views of
one flat buffer), across the 2-3 generator calls of a step, like autograd's accumulation does.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, cur_stream

_KEEP_SLOTS = 3      # x_global, y_global, x_entire may be alive at once; slot 3 serves no-grad calls


def pick_keep_slot(tokens) -> int:
    """Activation slot for a netG call whose backward is still to come. Lowest free slot first: the same call site
    gets the same slot every step (x_global -> 0, y_global -> 1, ...), which keeps the engine's (slot, shape)
    CUDA-graph keys few. When all are held, recycle the OLDEST pass: the one whose backward never came (step 0's
    y_global feeds no active loss term, util/losses.py:35-37), never a pass of the current step."""
    free = [i for i in range(_KEEP_SLOTS) if tokens[i] is None]
    if free:
        return free[0]
    return min(range(_KEEP_SLOTS), key=lambda i: tokens[i])


class _GenFn(torch.autograd.Function):
    """netG over one or several independent inputs on the native engine (one autograd node for all of them, so that
    the backward passes of a step's 2-3 netG calls can be issued together). `anchor` (a parameter) only makes
    autograd schedule backward(); the parameter gradients are written by the engine as a side effect, the way
    fused optimisers consume them."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, module: "NativeSkip", keep: bool, *xs: torch.Tensor):
        outs, slots, tokens = module._run_forward(xs, keep)
        ctx.module, ctx.slots, ctx.tokens = module, slots, tokens
        ctx.set_materialize_grads(False)   # an output no loss term consumed arrives as None: its backward is skipped
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        ctx.module._run_backward(gouts, ctx.slots, ctx.tokens)
        return (None, None, None) + (None,) * len(gouts)


class NativeSkip(nn.Sequential):
    """nn.Sequential tree of the default-argument skip() whose forward runs on the native generator engine."""

    def __init__(self):
        super().__init__()
        self._params_cache = None
        self._bns_cache = None
        self._eng = None
        self._ptrs = None
        self._ptr_sig = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._grad_views: List[torch.Tensor] = []
        self._slot_tokens = [None] * 4
        self._token = 0
        self._side_streams: List[torch.cuda.Stream] = []
        self._priv_grads: List[torch.Tensor] = []     # per-slot gradient buffers of concurrently issued backward passes
        self._priv_ptrs: dict = {}
        self.concurrent = True                        # independent netG calls of a step run on parallel streams

    def __getstate__(self):
        # copy.deepcopy / pickle: the copy owns its tensors but not this module's engine, pointer tables, streams or gradient buffers
        st = self.__dict__.copy()
        st.update(_params_cache=None, _bns_cache=None, _eng=None, _ptrs=None, _ptr_sig=None, _flat_grad=None, _grad_views=[],
                  _slot_tokens=[None] * 4, _side_streams=[], _priv_grads=[], _priv_ptrs={})
        return st

    # ---- pointer tables ----------------------------------------------------------------------------
    def _param_list(self) -> List[torch.nn.Parameter]:
        # cached: walking the module tree costs ~0.1 ms and this is asked for several times per step. The Parameter
        # OBJECTS of a module tree are stable (load_state_dict / .to() / optimizers update them in place); the cache is
        # dropped when a sub-module is added (add_module below).
        ps = self._params_cache
        if ps is None:
            ps = list(self.parameters())
            if len(ps) != _lib.GEN_PARAMS:
                raise RuntimeError(f"NativeSkip expects {_lib.GEN_PARAMS} parameter tensors, found {len(ps)}")
            self._params_cache = ps
        return ps

    def _bn_list(self) -> List[nn.BatchNorm2d]:
        bns = self._bns_cache
        if bns is None:
            bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
            if len(bns) != _lib.GEN_BN:
                raise RuntimeError(f"NativeSkip expects {_lib.GEN_BN} BatchNorm2d layers, found {len(bns)}")
            self._bns_cache = bns
        return bns

    def add_module(self, name, module):
        self.__dict__['_params_cache'] = None
        self.__dict__['_bns_cache'] = None
        return super().add_module(name, module)

    def _apply(self, fn, *a, **kw):
        # .to() / .cuda() / .float(): tensors may be replaced
        self.__dict__['_params_cache'] = None
        self.__dict__['_bns_cache'] = None
        self.__dict__['_ptr_sig'] = None
        return super()._apply(fn, *a, **kw)

    def _engine(self):
        if self._eng is None:
            h = C.c_void_p()
            check(_lib.splice_gen_create(C.byref(h)), "splice_gen_create")
            self._eng = h
        return self._eng

    def _refresh_ptrs(self, need_grads: bool):
        ps = self._param_list()
        sig = (ps[0].data_ptr(), ps[55].data_ptr(), ps[-1].data_ptr(), ps[0].device,
               None if not need_grads else (ps[0].grad.data_ptr() if ps[0].grad is not None else 0,
                                            ps[-1].grad.data_ptr() if ps[-1].grad is not None else 0))
        if self._ptrs is not None and sig == self._ptr_sig:
            return self._ptrs
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("the native generator needs contiguous fp32 CUDA parameters (no CPU fallback): "
                                   "move the model to a CUDA device")
        t = _lib.SpliceGenPointers()
        for i, p in enumerate(ps):
            t.param[i] = p.data_ptr()
            t.grad[i] = p.grad.data_ptr() if (need_grads and p.grad is not None) else None
        for i, bn in enumerate(self._bn_list()):
            t.running_mean[i] = bn.running_mean.data_ptr()
            t.running_var[i] = bn.running_var.data_ptr()
            t.num_batches_tracked[i] = bn.num_batches_tracked.data_ptr()
        self._ptrs, self._ptr_sig = t, sig
        return t

    def _attach_grads(self):
        """Make every p.grad a view into one flat buffer (allocated once); zero what was missing."""
        ps = self._param_list()
        if self._flat_grad is None or self._flat_grad.device != ps[0].device:
            n = sum(p.numel() for p in ps)
            self._flat_grad = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
            self._grad_views, off = [], 0
            for p in ps:
                self._grad_views.append(self._flat_grad[off:off + p.numel()].view_as(p))
                off += p.numel()
            for p, v in zip(ps, self._grad_views):
                p._splice_flat_grad, p._splice_grad_view = self._flat_grad, v
        for p, v in zip(ps, self._grad_views):   # a gradient tensor somebody else assigned: adopt its value, keep our view
            if p.grad is not None and p.grad is not v:
                v.copy_(p.grad)
                p.grad = v
                self._ptr_sig = None
        missing = [i for i, p in enumerate(ps) if p.grad is None]
        if len(missing) == len(ps):
            self._flat_grad.zero_()
            for p, v in zip(ps, self._grad_views):
                p.grad = v
            self._ptr_sig = None
        elif missing:
            for i in missing:
                self._grad_views[i].zero_()
                ps[i].grad = self._grad_views[i]
            self._ptr_sig = None

    # ---- execution ---------------------------------------------------------------------------------
    def _streams(self, n: int) -> List[torch.cuda.Stream]:
        while len(self._side_streams) < n:
            self._side_streams.append(torch.cuda.Stream())
        return self._side_streams[:n]

    def _private_table(self, slot: int):
        """Pointer table whose gradients point into this slot's private flat buffer (overwrite-mode backward)."""
        ps = self._param_list()
        key = (slot, ps[0].data_ptr(), ps[-1].data_ptr())
        hit = self._priv_ptrs.get(slot)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        while len(self._priv_grads) <= slot:
            self._priv_grads.append(torch.empty_like(self._flat_grad))
        flat = self._priv_grads[slot]
        t = _lib.SpliceGenPointers()
        off = 0
        for i, p in enumerate(ps):
            t.param[i] = p.data_ptr()
            t.grad[i] = flat.data_ptr() + 4 * off
            off += p.numel()
        self._priv_ptrs[slot] = (key, t, flat)
        return t, flat

    def _run_forward(self, xs, keep: bool):
        if not self.training:
            # nn.BatchNorm2d in eval mode normalises with the running statistics; the engine implements the training-mode
            # arithmetic only (the reference never calls .eval(), SURVEY hard part 6). Refuse rather than silently
            # returning batch-statistics images.
            raise NotImplementedError("splice_b200's native generator implements BatchNorm in training mode only (the "
                                      "reference never calls netG.eval()); call netG.train() before using it")
        prepared = []
        for x in xs:
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError("netG expects [N,3,H,W]")
            x = x.detach()
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.float().contiguous()
            if not x.is_cuda:
                raise RuntimeError("the native generator runs on sm_100a only; there is no CPU fallback")
            prepared.append(x)
        if not keep and len(prepared) > 1:
            raise RuntimeError("several no-grad netG calls share one activation slot: call them one at a time")
        t = self._refresh_ptrs(need_grads=False)
        main = torch.cuda.current_stream()
        side = self._streams(len(prepared)) if (self.concurrent and len(prepared) > 1) else None
        outs, slots, tokens = [], [], []
        for k, x in enumerate(prepared):
            n, _, h, w = x.shape
            out = torch.empty_like(x)
            slot = pick_keep_slot(self._slot_tokens) if keep else 3
            self._token += 1
            self._slot_tokens[slot] = self._token
            if side is not None:
                side[k].wait_stream(main)
                stream = side[k].cuda_stream
            else:
                stream = main.cuda_stream
            update_now = self.training and side is None
            check(_lib.splice_gen_forward(self._engine(), C.byref(t), x.data_ptr(), n, h, w, out.data_ptr(), slot,
                                          1 if keep else 0, 1 if update_now else 0, stream), "splice_gen_forward")
            outs.append(out); slots.append(slot); tokens.append(self._token)
        if side is not None:
            for st in side:
                main.wait_stream(st)
            if self.training:   # running statistics: applied in call order, like the reference's sequential calls
                for slot in slots:
                    check(_lib.splice_gen_update_running(self._engine(), C.byref(t), slot, main.cuda_stream),
                          "splice_gen_update_running")
        return outs, slots, tokens

    def _run_backward(self, gouts, slots, tokens):
        live = []
        for gout, slot, token in zip(gouts, slots, tokens):
            if self._slot_tokens[slot] != token:
                raise RuntimeError("generator activations were overwritten: more than 3 netG calls were kept alive "
                                   "before their backward pass")
            self._slot_tokens[slot] = None
            if gout is None:
                continue
            gout = gout.detach()
            if gout.dtype != torch.float32 or not gout.is_contiguous():
                gout = gout.float().contiguous()
            live.append((gout, slot))
        if not live:
            return
        self._attach_grads()
        main = torch.cuda.current_stream()
        if len(live) == 1 or not self.concurrent:
            t = self._refresh_ptrs(need_grads=True)
            for gout, slot in live:
                check(_lib.splice_gen_backward(self._engine(), C.byref(t), gout.data_ptr(), slot, 1, main.cuda_stream),
                      "splice_gen_backward")
            return
        # independent passes: each on its own stream into its own gradient buffer, then one fold into .grad
        side = self._streams(len(live))
        privs = []
        for (gout, slot), st in zip(live, side):
            t, flat = self._private_table(slot)
            st.wait_stream(main)
            check(_lib.splice_gen_backward(self._engine(), C.byref(t), gout.data_ptr(), slot, 0, st.cuda_stream),
                  "splice_gen_backward")
            privs.append(flat)
        for st in side:
            main.wait_stream(st)
        srcs = (C.c_void_p * len(privs))(*[f.data_ptr() for f in privs])
        check(_lib.splice_accumulate(self._flat_grad.data_ptr(), srcs, len(privs), self._flat_grad.numel(), main.cuda_stream),
              "splice_accumulate")

    def forward_many(self, inputs):
        """[netG(x) for x in inputs] — the independent netG calls of one step (models/model.py:16-23), issued together."""
        inputs = list(inputs)
        anchor = next(self.parameters())
        keep = torch.is_grad_enabled() and anchor.requires_grad   # (grad mode is off inside Function.forward)
        if not keep:
            return [_GenFn.apply(anchor, self, False, x)[0] for x in inputs]
        return list(_GenFn.apply(anchor, self, True, *inputs))

    def forward(self, input):
        return self.forward_many([input])[0]


    def __del__(self):
        eng = getattr(self, "_eng", None)
        if eng:
            try:
                torch.cuda.synchronize()
                _lib.splice_gen_destroy(eng)
            except Exception:  # noqa: BLE001
                pass
