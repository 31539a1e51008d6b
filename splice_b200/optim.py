"""Fused multi-tensor Adam behind the torch.optim.Optimizer interface.

`get_optimizer` (util/util.py) returns this for cfg['optimizer'] == 'adam'. Same defaults, `param_groups`
keys and per-parameter state names ('step', 'exp_avg', 'exp_avg_sq') as torch.optim.Adam, so
`optimizer.param_groups[0]['lr']` (train.py:64), LR schedulers (util.py:22) and `state_dict()` keep working.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, cur_stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam implements the configuration Splice uses: no weight decay, no amsgrad")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=True)
        super().__init__(params, defaults)
        self._fast = {}   # param-group index -> cached pointer tables + common step count (see step())

    def _flush_fast(self, gi=None):
        """Write the cached common step count back into the per-parameter state (before anything reads it)."""
        for k in ([gi] if gi is not None else list(self._fast)):
            c = self._fast.pop(k, None)
            if c is not None:
                for p in c['params']:
                    self.state[p]['_step'] = c['step']

    def load_state_dict(self, state_dict):
        self._fast.clear()
        return super().load_state_dict(state_dict)

    def zero_grad(self, set_to_none: bool = True):
        """Gradients that live in the native generator's flat buffer are cleared with ONE memset and stay attached
        (so the engine's pointer table stays valid); anything else follows torch's zero_grad."""
        flats, rest = {}, []
        for group in self.param_groups:
            for p in group['params']:
                flat = getattr(p, '_splice_flat_grad', None)
                if flat is not None and p.grad is not None and p.grad is getattr(p, '_splice_grad_view', None):
                    flats[id(flat)] = flat
                else:
                    rest.append(p)
        for flat in flats.values():
            flat.zero_()
        for p in rest:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            # steady state: same parameters, same (flat-buffer) gradient addresses, one common step count -> the ctypes
            # pointer tables of the previous step are reused and the per-parameter bookkeeping is one integer
            # (the key also covers the parameter and state addresses at both ends of the list: `.to()` / state surgery
            # that moves them while the flat gradient stays put must not leave stale pointers in the tables)
            def _ends(p):
                st = self.state.get(p, {})
                return (p.data_ptr(), st['exp_avg'].data_ptr() if 'exp_avg' in st else 0,
                        st['exp_avg_sq'].data_ptr() if 'exp_avg_sq' in st else 0)
            key = tuple(p.grad.data_ptr() for p in ps) + _ends(ps[0]) + _ends(ps[-1])
            cached = self._fast.get(gi)
            if cached is not None and cached['key'] == key and all(p.grad.is_contiguous() for p in ps):
                cached['step'] += 1
                check(_lib.splice_adam_step(*cached['tables'], cached['n'], cached['step'], float(group['lr']),
                                            float(group['betas'][0]), float(group['betas'][1]), float(group['eps']),
                                            cur_stream()), "splice_adam_step")
                continue
            self._flush_fast(gi)
            by_step = {}
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['_step'] = 0
                elif '_step' not in st:  # state restored by load_state_dict
                    st['_step'] = int(st['step'].item())
                st['_step'] += 1
                by_step.setdefault(st['_step'], []).append(p)
            # parameters that joined at different times need different bias corrections: one call per step count
            for step, sel in by_step.items():
                n = len(sel)
                arr = lambda xs: (C.c_void_p * n)(*xs)  # noqa: E731
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in sel]
                tables = (arr([p.data_ptr() for p in sel]), arr([g.data_ptr() for g in grads]),
                          arr([self.state[p]['exp_avg'].data_ptr() for p in sel]),
                          arr([self.state[p]['exp_avg_sq'].data_ptr() for p in sel]),
                          (C.c_int * n)(*[p.numel() for p in sel]))
                check(_lib.splice_adam_step(*tables, n, step, float(group['lr']),
                                            float(group['betas'][0]), float(group['betas'][1]), float(group['eps']),
                                            cur_stream()), "splice_adam_step")
                if len(by_step) == 1 and len(sel) == len(ps) and all(g is p.grad for g, p in zip(grads, sel)):
                    key = tuple(p.grad.data_ptr() for p in ps) + _ends(ps[0]) + _ends(ps[-1])   # state exists now
                    self._fast[gi] = {'key': key, 'tables': tables, 'n': n, 'step': step, 'params': list(sel)}
        return loss

    def state_dict(self):
        # the per-parameter 'step' tensors torch.optim.Adam exposes are materialised lazily (host-side bookkeeping
        # uses a python int so that a training step does not touch 112 CPU tensors)
        self._flush_fast()
        for st in self.state.values():
            if '_step' in st:
                st['step'].fill_(float(st['_step']))
        sd = super().state_dict()
        for st in sd['state'].values():
            st.pop('_step', None)
        return sd
