"""ORACLE tooling: golden vectors of inversion.py's generator, made by the UNMODIFIED reference (models/unet/skip.py imported from
/root/reference; runs only in the build container).

    python -m oracle.make_golden_inversion      # (re)writes tests/golden/inversion_gen.pt

The reference builds `skip(32, 3, [16,32,64,128,128,128], ..., filter_size_down = filter_size_up = [7,7,5,5,3,3], 'stride',
pad='reflection')` (inversion.py:21-25), keeps torch's default initialisation (no init_weights) and calls `net(net_input)` in training
mode. Pinned here, all on the CPU in float32 under fixed seeds:
  keys / shapes / fingerprints of the freshly constructed state_dict (seed 0): module naming and RNG order of the construction;
  y = net(x) for a seeded 1x32x72x104 noise input (regenerated from its seed by the tests, not stored), after perturbing BatchNorm affine parameters with a seeded recipe;
  fingerprints of every parameter gradient of (y * w).sum(), and of the BatchNorm running statistics after that pass.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"

INVERSION_ARGS = dict(num_channels_down=[16, 32, 64, 128, 128, 128], num_channels_up=[16, 32, 64, 128, 128, 128],
                      num_channels_skip=[4, 4, 4, 4, 4, 4], filter_size_down=[7, 7, 5, 5, 3, 3], filter_size_up=[7, 7, 5, 5, 3, 3],
                      downsample_mode='stride', pad='reflection')


def fingerprint(t: torch.Tensor) -> dict:
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(16, f.numel())).long()
    return {"shape": tuple(t.shape), "sum": f.sum().item(), "abs": f.abs().sum().item(), "samples": f[idx].clone()}


def perturb(net, seed: int) -> None:
    """Same recipe as tools/genx_compare.randomise (kept separate: the oracle does not import the tools)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(1.0 + 0.3 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, torch.nn.Conv2d):
                m.weight.copy_(m.weight + 0.05 * torch.randn(m.weight.shape, generator=g))


def golden_input():
    """(x, w): the seeded noise input and the seeded cotangent of the golden pass (CPU generators: machine-independent)."""
    return (torch.randn(1, 32, 72, 104, generator=torch.Generator().manual_seed(2)),
            torch.randn(1, 3, 72, 104, generator=torch.Generator().manual_seed(3)))


def run(skip_fn) -> dict:
    """Builds the network with `skip_fn` and evaluates it with torch's modules (`nn.Sequential.forward`)."""
    torch.manual_seed(0)
    net = skip_fn(32, 3, **INVERSION_ARGS)
    out = {"keys": list(net.state_dict().keys()), "init": {k: fingerprint(v) for k, v in net.state_dict().items()}}
    perturb(net, 1)
    x, w = golden_input()
    y = torch.nn.Sequential.forward(net, x)
    (y * w).sum().backward()
    out.update(y=y.detach().clone(), grads={k: fingerprint(p.grad) for k, p in net.named_parameters()},
               buffers={k: fingerprint(v) for k, v in net.state_dict().items() if "running" in k or "tracked" in k})
    return out


def main() -> None:
    sys.path.insert(0, str(REF))
    from models.unet.skip import skip as ref_skip   # the reference's own builder

    gold = run(ref_skip)
    GOLD.mkdir(exist_ok=True)
    torch.save(gold, GOLD / "inversion_gen.pt")
    print("wrote", GOLD / "inversion_gen.pt", "keys", len(gold["keys"]), "y mean", gold["y"].mean().item())


if __name__ == "__main__":
    main()
