"""Comparison of a NativeSkipX network (csrc/generator_x.cu) with torch evaluating the same module tree - shared by the CPU
emulation tests (tests/test_genx_emu.py) and the GPU check (tools/gpu_checks.py: generator_inversion_variant).

Reference = torch in float64 on a deep copy of the tree (`nn.Sequential.forward`, i.e. nn.ReflectionPad2d / Conv2d / BatchNorm2d /
LeakyReLU / Upsample / Concat module by module, what the reference's skip() does) and its autograd; yardstick = the same in float32.
Two things make "max error below a fixed epsilon" the wrong bar for these networks:
  * BatchNorm over a handful of pixels at the deepest scales (6 scales: 224 px -> 4 x 4) amplifies rounding differences: torch's own
    float32 path is 1e-5 .. 5e-2 away from float64 depending on the layer, so errors are judged against that distance;
  * LeakyReLU' is discontinuous: a pre-activation within rounding of zero (|z| ~ 1e-6: a few among 10^5..10^6 values) gets the other
    slope in one of the two evaluations and shifts that channel's gradient by 0.8 x one pixel's share. `strict=True` is for small
    networks on inputs without such near-ties (see `tie_margin`); otherwise a flip allowance is added and a global L2 bound is
    checked besides.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn


def randomise(model, seed):
    """Non-trivial BatchNorm affine parameters and perturbed conv weights, as after a few optimisation steps."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.copy_((1.0 + 0.3 * torch.randn(m.weight.shape, generator=g)).to(m.weight.device))
                m.bias.copy_((0.2 * torch.randn(m.bias.shape, generator=g)).to(m.bias.device))
            elif isinstance(m, nn.Conv2d):
                m.weight.copy_(m.weight + (0.05 * torch.randn(m.weight.shape, generator=g)).to(m.weight.device))


def reference_pass(model, x, w):
    """torch evaluates the tree module by module; returns output, gradients, BatchNorm buffers after the pass."""
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    for p in model.parameters():
        p.grad = None
    y = nn.Sequential.forward(model, x)
    (y * w).sum().backward()
    return (y.detach(), [p.grad.clone() for p in model.parameters()], [b.running_mean.clone() for b in bns],
            [b.running_var.clone() for b in bns], [int(b.num_batches_tracked) for b in bns])


def tie_margin(model, x):
    """Smallest |z| over every LeakyReLU input of the float64 evaluation: inputs for `strict` comparisons are chosen with a margin
    well above float32 rounding, so that both evaluations take the same slope everywhere."""
    m64 = copy.deepcopy(model).double()
    lo = [float("inf")]
    hooks = [m.register_forward_pre_hook(lambda mod, inp: lo.__setitem__(0, min(lo[0], inp[0].abs().min().item())))
             for m in m64.modules() if isinstance(m, nn.LeakyReLU)]
    with torch.no_grad():
        nn.Sequential.forward(m64, x.double())
    for h in hooks:
        h.remove()
    return lo[0]


def compare(model, x, seed, strict=False):
    """Runs model(x) and backward on the native engine, the float64 / float32 references on copies; asserts and returns a summary.
    Gradients already present on `model` are accumulated into (autograd semantics) and that is checked as such."""
    dev = x.device
    m64 = copy.deepcopy(model).double()
    m32 = copy.deepcopy(model)
    w = torch.randn((x.shape[0], model._config["num_output_channels"], x.shape[2], x.shape[3]),
                    generator=torch.Generator().manual_seed(seed)).to(dev)
    y64, g64, rm64, rv64, nbt64 = reference_pass(m64, x.double(), w.double())
    y32, g32, _, _, _ = reference_pass(m32, x, w)
    had = [None if p.grad is None else p.grad.clone() for p in model.parameters()]

    y = model(x)
    assert y.shape == y64.shape and torch.isfinite(y).all()
    fwd_yard = (y32.double() - y64).abs().max().item()
    fwd_err = (y.double() - y64).abs().max().item()
    assert fwd_err <= 4 * fwd_yard + 2e-6, ("forward", fwd_err, fwd_yard)
    (y * w).sum().backward()

    flip = 0.0 if strict else 3e-2
    worst, num, den, ynum = 0.0, 0.0, 0.0, 0.0
    # conv biases in front of a BatchNorm have an exactly-zero gradient; what both evaluations return for them is cancellation noise
    # proportional to the gradients around them, hence a floor relative to the largest gradient of the network
    floor = max(1e-3, 1e-3 * max(a.abs().max().item() for a in g64))
    for i, (p, a, b, prev) in enumerate(zip(model.parameters(), g64, g32, had)):
        assert p.grad is not None and torch.isfinite(p.grad).all(), i
        mine = p.grad.double() - (0 if prev is None else prev.double())
        scale = max(a.abs().max().item(), floor)
        yard = (b.double() - a).abs().max().item() / scale
        err = (mine - a).abs().max().item() / scale
        assert err <= 4 * yard + 2e-5 + flip, ("gradient", i, tuple(p.shape), err, yard)
        worst = max(worst, err)
        num += (mine - a).pow(2).sum().item(); ynum += (b.double() - a).pow(2).sum().item(); den += a.pow(2).sum().item()
    l2, l2_yard = (num / den) ** 0.5, (ynum / den) ** 0.5
    assert l2 <= 4 * l2_yard + (1e-6 if strict else 5e-3), ("gradient L2", l2, l2_yard)
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    for b, m, v, n in zip(bns, rm64, rv64, nbt64):
        assert torch.allclose(b.running_mean.double(), m, atol=1e-5, rtol=1e-4)
        assert torch.allclose(b.running_var.double(), v, atol=1e-5, rtol=1e-4)
        assert int(b.num_batches_tracked) == n
    return {"fwd_err": fwd_err, "fwd_yard": fwd_yard, "grad_worst_rel": worst, "grad_l2_rel": l2, "grad_l2_yard": l2_yard}
