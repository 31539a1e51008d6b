"""Host-side glue for skip() configurations other than the optimisation loop's default (include/splice_b200.h:
splice_genx_*): inversion.py:21-25 builds a 6-scale, 32-input-channel, 7x7 / 5x5 / 3x3, reflection-padded network and
runs `net(net_input)` / `loss.backward()` / `optimizer.step()` on it once per iteration (inversion.py:65-69).

Like `NativeSkip`, the nn.Sequential tree built by models/unet/skip.py only OWNS the tensors (same state_dict keys,
`parameters()` order and initial weights as the reference); `forward` sends them by pointer to the native engine, whose
backward writes the parameter gradients straight into `.grad` (views of one flat buffer). One pass is kept for backward
at a time - the inversion loop has exactly one netG call per iteration.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib

# where the engine lives. The product value is the CUDA library; tests/test_genx_emu.py swaps in the CPU emulation build of the
# SAME sources (tests/emu) to exercise this glue without a GPU. Nothing in the package ever assigns it.
_backend = _lib


def _stream() -> int:
    return _backend.cur_stream()


class _GenXFn(torch.autograd.Function):
    """net(x) on the native engine. `anchor` (a parameter) only makes autograd schedule backward(); the parameter gradients
    are written by the engine as a side effect, the way fused optimisers consume them."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, module: "NativeSkipX", keep: bool, x: torch.Tensor):
        out, token = module._run_forward(x, keep)
        ctx.module, ctx.token = module, token
        return out

    @staticmethod
    def backward(ctx, gout):
        ctx.module._run_backward(gout, ctx.token)
        return None, None, None, None


class NativeSkipX(nn.Sequential):
    """nn.Sequential tree of a non-default skip() whose forward runs on the native generalised generator engine."""

    def __init__(self, config: dict):
        super().__init__()
        self._config = dict(config)
        self._params_cache = None
        self._bns_cache = None
        self._eng = None
        self._bound_sig = None
        self._bound_gsig = None
        self._tables = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._grad_views: List[torch.Tensor] = []
        self._token = 0
        self._kept_token = None

    def __getstate__(self):
        # copy.deepcopy / pickle: the copy owns its tensors but not this module's engine, pointer tables or gradient buffer
        st = self.__dict__.copy()
        st.update(_params_cache=None, _bns_cache=None, _eng=None, _bound_sig=None, _bound_gsig=None, _tables=None, _flat_grad=None,
                  _grad_views=[], _kept_token=None)
        return st

    # ---- tables --------------------------------------------------------------------------------------
    def _expected_counts(self):
        n = len(self._config["num_channels_down"])
        return 22 * n + 2, 6 * n

    def _param_list(self) -> List[torch.nn.Parameter]:
        ps = self._params_cache
        if ps is None:
            ps = list(self.parameters())
            if len(ps) != self._expected_counts()[0]:
                raise RuntimeError(f"NativeSkipX expects {self._expected_counts()[0]} parameter tensors, found {len(ps)}")
            self._params_cache = ps
        return ps

    def _bn_list(self) -> List[nn.BatchNorm2d]:
        bns = self._bns_cache
        if bns is None:
            bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
            if len(bns) != self._expected_counts()[1]:
                raise RuntimeError(f"NativeSkipX expects {self._expected_counts()[1]} BatchNorm2d layers, found {len(bns)}")
            self._bns_cache = bns
        return bns

    def add_module(self, name, module):
        self.__dict__['_params_cache'] = None
        self.__dict__['_bns_cache'] = None
        return super().add_module(name, module)

    def _apply(self, fn, *a, **kw):
        self.__dict__['_params_cache'] = None
        self.__dict__['_bns_cache'] = None
        self.__dict__['_bound_sig'] = None
        self.__dict__['_bound_gsig'] = None
        return super()._apply(fn, *a, **kw)

    def _engine(self):
        if self._eng is None:
            cfg = self._config
            n = len(cfg["num_channels_down"])
            c = _lib.SpliceGenXConfig()
            c.n_scales = n
            c.in_channels, c.out_channels = cfg["num_input_channels"], cfg["num_output_channels"]
            for i in range(n):
                c.ch_down[i], c.ch_up[i], c.ch_skip[i] = cfg["num_channels_down"][i], cfg["num_channels_up"][i], cfg["num_channels_skip"][i]
                c.k_down[i], c.k_up[i] = cfg["filter_size_down"][i], cfg["filter_size_up"][i]
            c.k_skip = cfg["filter_skip_size"]
            c.reflect = 1 if cfg["pad"] == "reflection" else 0
            c.sigmoid = 1 if cfg["need_sigmoid"] else 0
            h = C.c_void_p()
            _backend.check(_backend.splice_genx_create(C.byref(c), C.byref(h)), "splice_genx_create")
            n_params, n_bn = C.c_int(), C.c_int()
            _backend.check(_backend.splice_genx_counts(h, C.byref(n_params), C.byref(n_bn)), "splice_genx_counts")
            if (n_params.value, n_bn.value) != self._expected_counts():
                raise RuntimeError("engine / module tree disagree on the number of tensors")
            self._eng = h
        return self._eng

    def _check_tensor(self, t: torch.Tensor, what: str):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"the native generator needs contiguous fp32 {what}")
        if _backend is _lib and not t.is_cuda:
            raise RuntimeError("the native generator runs on sm_100a only; there is no CPU fallback: move the model and its "
                               "input to a CUDA device")

    def _bind(self, need_grads: bool):
        """(Re)send the pointer tables when a tensor has moved. A forward-only bind leaves the gradient table alone."""
        ps, bns = self._param_list(), self._bn_list()
        psig = (tuple(p.data_ptr() for p in ps), tuple(b.running_mean.data_ptr() for b in bns))
        have_grads = all(p.grad is not None for p in ps)
        if need_grads and not have_grads:
            raise RuntimeError("gradient buffers are not attached")
        gsig = tuple(p.grad.data_ptr() for p in ps) if have_grads else None
        if psig == self._bound_sig and (gsig == self._bound_gsig or not need_grads):
            return
        for p in ps:
            self._check_tensor(p.data, "parameters")
        n, nb = len(ps), len(bns)
        params = (C.c_void_p * n)(*[p.data_ptr() for p in ps])
        grads = (C.c_void_p * n)(*[p.grad.data_ptr() for p in ps]) if have_grads else None
        rmean = (C.c_void_p * nb)(*[b.running_mean.data_ptr() for b in bns])
        rvar = (C.c_void_p * nb)(*[b.running_var.data_ptr() for b in bns])
        nbt = (C.c_void_p * nb)(*[b.num_batches_tracked.data_ptr() for b in bns])
        _backend.check(_backend.splice_genx_bind(self._engine(), params, grads, rmean, rvar, nbt), "splice_genx_bind")
        self._tables = (params, grads, rmean, rvar, nbt)
        self._bound_sig, self._bound_gsig = psig, gsig

    def _attach_grads(self):
        """Make every p.grad a view into one flat buffer (allocated once). Returns whether the engine has to ACCUMULATE
        (some gradient already held a value) or may overwrite (all were None)."""
        ps = self._param_list()
        if self._flat_grad is None or self._flat_grad.device != ps[0].device:
            n = sum(p.numel() for p in ps)
            self._flat_grad = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
            self._grad_views, off = [], 0
            for p in ps:
                self._grad_views.append(self._flat_grad[off:off + p.numel()].view_as(p))
                off += p.numel()
            for p, v in zip(ps, self._grad_views):   # optim.FusedAdam recognises parameters whose gradients share one buffer
                p._splice_flat_grad, p._splice_grad_view = self._flat_grad, v
        if all(p.grad is None for p in ps):      # optimizer.zero_grad() sets .grad to None (inversion.py:64): nothing to add to,
            for p, v in zip(ps, self._grad_views):   # the engine overwrites every element
                p.grad = v
            return False
        for p, v in zip(ps, self._grad_views):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad is not v:                # a gradient tensor somebody else assigned: adopt its value, keep our view
                v.copy_(p.grad)
                p.grad = v
        return True

    # ---- execution -----------------------------------------------------------------------------------
    def _run_forward(self, x: torch.Tensor, keep: bool):
        if not self.training:
            raise NotImplementedError("splice_b200's native generator implements BatchNorm in training mode only (the "
                                      "reference never calls net.eval()); call net.train() before using it")
        cin = self._config["num_input_channels"]
        if x.dim() != 4 or x.shape[1] != cin:
            raise ValueError(f"net expects [N,{cin},H,W]")
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        self._check_tensor(x, "input")
        self._bind(need_grads=False)
        n, _, h, w = x.shape
        out = torch.empty((n, self._config["num_output_channels"], h, w), device=x.device, dtype=torch.float32)
        _backend.check(_backend.splice_genx_forward(self._engine(), x.data_ptr(), n, h, w, out.data_ptr(), 1 if keep else 0, 1, _stream()),
                       "splice_genx_forward")
        self._token += 1
        self._kept_token = self._token if keep else None
        return out, self._token

    def _run_backward(self, gout: torch.Tensor, token: int):
        if self._kept_token != token:
            raise RuntimeError("generator activations were overwritten: a later net(x) call replaced the pass this gradient belongs to")
        self._kept_token = None
        gout = gout.detach()
        if gout.dtype != torch.float32 or not gout.is_contiguous():
            gout = gout.float().contiguous()
        accumulate = self._attach_grads()
        self._bind(need_grads=True)
        _backend.check(_backend.splice_genx_backward(self._engine(), gout.data_ptr(), 1 if accumulate else 0, _stream()),
                       "splice_genx_backward")

    def forward(self, input):
        if torch.is_grad_enabled() and input.requires_grad:
            # the engine stops at the parameter gradients (what inversion.py's optimiser consumes, inversion.py:50,68); returning
            # None for d loss / d input silently would be a wrong answer, so say so
            raise NotImplementedError("the native generator does not compute the gradient with respect to its input: pass "
                                      "net_input.detach() (inversion.py optimises the network's parameters only)")
        anchor = next(self.parameters())
        keep = torch.is_grad_enabled() and anchor.requires_grad
        return _GenXFn.apply(anchor, self, keep, input)

    def __del__(self):
        eng = getattr(self, "_eng", None)
        if eng:
            try:
                if _backend is _lib:
                    torch.cuda.synchronize()
                _backend.splice_genx_destroy(eng)
            except Exception:  # noqa: BLE001
                pass
