"""The generalised generator engine (csrc/generator_x.cu: inversion.py's 6-scale, 7x7 / 5x5, reflection-padded skip() and
other non-default configurations) checked WITHOUT a GPU: the same source is compiled by g++ against tests/emu/cuda_emu.h
(kernel bodies run as CPU fibers, host orchestration unchanged) and driven through the product's own Python glue
(splice_b200/generator_x.py), then compared with torch evaluating the module tree the reference would build
(nn.ReflectionPad2d / Conv2d / BatchNorm2d / LeakyReLU / Upsample / Concat) and its autograd.

This is test infrastructure: the product library is the nvcc build and has no CPU path; the `-m gpu` check
`generator_inversion_variant` repeats the comparison on the real kernels.
"""
import copy
import ctypes as C
import sys
from pathlib import Path

import pytest
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


class _EmuBackend:
    """The emulation build behind the names generator_x.py uses on its backend."""

    def __init__(self):
        sys.path.insert(0, str(ROOT / "tests" / "emu"))
        import build_emu
        from splice_b200 import _lib

        self.lib = C.CDLL(str(build_emu.build()))
        for name, (restype, argtypes) in _lib.GENX_SIGNATURES.items():
            fn = getattr(self.lib, name)
            fn.restype, fn.argtypes = restype, argtypes
            setattr(self, name, fn)
        self._err = _lib.SpliceError

    def check(self, rc, what=""):
        if rc != 0:
            raise self._err(f"{what} failed in the emulation build (rc={rc})")

    @staticmethod
    def cur_stream():
        return None


@pytest.fixture(scope="module")
def emu_backend():
    from splice_b200 import generator_x

    be = _EmuBackend()
    old = generator_x._backend
    generator_x._backend = be
    yield be
    generator_x._backend = old


def _tie_free_input(net, shape, first_seed, margin=2e-5):
    """A seeded input on which no LeakyReLU pre-activation of the float64 evaluation is within `margin` of zero."""
    from tools.genx_compare import tie_margin

    for seed in range(first_seed, first_seed + 200):
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
        if tie_margin(net, x) > margin:
            return x
    raise AssertionError("no tie-free input found")


def test_inversion_variant_matches_torch_modules(emu_backend):
    """inversion.py:21-25's network (6 scales, 7x7 / 5x5 / 3x3, reflection padding, 32-channel noise input) at a small, odd,
    non-square size: forward, every parameter gradient, running statistics (the loop's later iterations - zero_grad(), noise added
    to the input, graph replay - are covered on the GPU by generator_inversion_variant's adam_loop part)."""
    from splice_b200.generator_x import NativeSkipX
    from splice_b200.models.unet.skip import skip
    from tools.genx_compare import compare, randomise

    torch.manual_seed(0)
    net = skip(32, 3, num_channels_down=[16, 32, 64, 128, 128, 128], num_channels_up=[16, 32, 64, 128, 128, 128],
               num_channels_skip=[4, 4, 4, 4, 4, 4], filter_size_down=[7, 7, 5, 5, 3, 3], filter_size_up=[7, 7, 5, 5, 3, 3],
               downsample_mode='stride', pad='reflection')
    assert isinstance(net, NativeSkipX)
    assert len(list(net.parameters())) == 134
    randomise(net, 1)
    x = torch.randn(1, 32, 67, 90, generator=torch.Generator().manual_seed(2))
    compare(net, x, 3)
    with pytest.raises(NotImplementedError, match="gradient with respect to its input"):
        net(x.clone().requires_grad_(True))


@pytest.mark.parametrize("pad", ["zero", "reflection"])
def test_three_scales_strict(emu_backend, pad):
    """Three scales of the inversion network (7x7, 7x7, 5x5; odd sizes so that strided and cropped borders are exercised) on an input
    without LeakyReLU near-ties: every gradient as close to float64 as torch's float32 path, no allowance."""
    from splice_b200.models.unet.skip import skip
    from tools.genx_compare import compare, randomise

    torch.manual_seed(20)
    net = skip(32, 3, num_channels_down=[16, 32, 64], num_channels_up=[16, 32, 64], num_channels_skip=[4, 4, 4],
               filter_size_down=[7, 7, 5], filter_size_up=[7, 7, 5], pad=pad)
    randomise(net, 21)
    x = _tie_free_input(net, (1, 32, 43, 58), 22)
    compare(net, x, 23, strict=True)


@pytest.mark.parametrize("pad", ["zero", "reflection"])
def test_small_mixed_configuration(emu_backend, pad):
    """A 2-scale network with batch 2, 3x3 skip filters, mixed filter sizes and no sigmoid, so that the batch dimension, the
    accumulate path and the linear output of the generalised kernels are exercised too."""
    from splice_b200.models.unet.skip import skip
    from tools.genx_compare import compare, randomise

    torch.manual_seed(10)
    net = skip(5, 2, num_channels_down=[8, 12], num_channels_up=[8, 12], num_channels_skip=[3, 5], filter_size_down=[5, 3],
               filter_size_up=[3, 7], filter_skip_size=3, need_sigmoid=False, pad=pad)
    randomise(net, 11)
    x = _tie_free_input(net, (2, 5, 21, 30), 12)
    compare(net, x, 13, strict=True)
    # gradients left in place are accumulated into (autograd semantics): the same comparison again without clearing them
    assert all(p.grad is not None for p in net.parameters())
    compare(net, x, 14, strict=True)


def test_unsupported_configurations_are_refused():
    from splice_b200.models.unet.skip import skip

    with pytest.raises(NotImplementedError):
        skip(3, 3, upsample_mode='nearest')
    with pytest.raises(NotImplementedError):
        skip(3, 3, act_fun='Swish')
    with pytest.raises(NotImplementedError):
        skip(3, 3, filter_size_down=9)
    with pytest.raises(NotImplementedError):
        skip(3, 3, downsample_mode='avg')


# ---- against the reference's own builder (golden made by oracle/make_golden_inversion.py from /root/reference) ------------
def _load_inversion_golden():
    return torch.load(ROOT / "tests" / "golden" / "inversion_gen.pt", weights_only=False)


def _close(fp, t, rel):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(16, f.numel())).long()
    tol = rel * max(fp["abs"] / max(f.numel(), 1), 1e-6)
    return (tuple(t.shape) == tuple(fp["shape"]) and abs(f.sum().item() - fp["sum"]) <= rel * max(fp["abs"], 1e-6)
            and (f[idx] - fp["samples"]).abs().max().item() <= 50 * tol)


def test_inversion_tree_and_init_match_reference():
    """Same state_dict keys (module naming incl. the ReflectionPad2d children), same shapes and - under the same seed - the same
    default initialisation as the reference's skip() (construction order = RNG order)."""
    from oracle.make_golden_inversion import INVERSION_ARGS
    from splice_b200.inversion import NET_ARGS
    from splice_b200.models.unet.skip import skip

    assert NET_ARGS == INVERSION_ARGS      # what the package's inversion script builds == what the golden was made with
    gold = _load_inversion_golden()
    torch.manual_seed(0)
    net = skip(32, 3, **NET_ARGS)
    sd = net.state_dict()
    assert list(sd.keys()) == gold["keys"]
    for k, v in sd.items():
        fp = gold["init"][k]
        f = v.reshape(-1).double()
        idx = torch.linspace(0, f.numel() - 1, min(16, f.numel())).long()
        assert tuple(v.shape) == tuple(fp["shape"]) and torch.equal(f[idx], fp["samples"]) and f.sum().item() == fp["sum"], k


def test_inversion_engine_matches_reference_golden(emu_backend):
    """The (emulated) engine on the golden input: output against the reference's own y, every parameter gradient and BatchNorm
    buffer against the reference's fingerprints (tolerances: float32 through 36 BatchNorm layers, see tools/genx_compare.py)."""
    from oracle.make_golden_inversion import INVERSION_ARGS, golden_input, perturb
    from splice_b200.models.unet.skip import skip

    gold = _load_inversion_golden()
    torch.manual_seed(0)
    net = skip(32, 3, **INVERSION_ARGS)
    perturb(net, 1)
    x, w = golden_input()
    y = net(x)
    assert (y - gold["y"]).abs().max().item() < 5e-4
    (y * w).sum().backward()
    # conv biases in front of a BatchNorm have an exactly-zero gradient (both sides return cancellation noise for them)
    numel = {k: p.numel() for k, p in net.named_parameters()}
    big = max(fp["abs"] / numel[k] for k, fp in gold["grads"].items())
    bad = []
    for k, p in net.named_parameters():
        fp = gold["grads"][k]
        if fp["abs"] / numel[k] < 1e-4 * big:
            ok = p.grad.abs().mean().item() < 1e-3 * big
        else:
            ok = _close(fp, p.grad, 5e-2)
        if not ok:
            bad.append(k)
    assert not bad, bad
    bad = [k for k, v in net.state_dict().items() if k in gold["buffers"] and not _close(gold["buffers"][k], v.float(), 1e-3)]
    assert not bad, bad


def test_glue_lifecycle(emu_backend):
    """NativeSkipX's host glue on the emulated engine: deepcopy gives an independent module with its own engine, load_state_dict /
    in-place parameter updates are seen by the next pass, a foreign .grad tensor is adopted (and accumulated into), a pass replaced
    by a later forward refuses its backward, no_grad forwards keep nothing."""
    from splice_b200.models.unet.skip import skip
    from tools.genx_compare import randomise

    torch.manual_seed(30)
    net = skip(4, 3, num_channels_down=[8, 8], num_channels_up=[8, 8], num_channels_skip=[2, 2], filter_size_down=[3, 5],
               filter_size_up=[5, 3], pad="reflection")
    randomise(net, 31)
    x = torch.randn(1, 4, 19, 26, generator=torch.Generator().manual_seed(32))
    ref = lambda m: nn.Sequential.forward(copy.deepcopy(m), x).detach()      # noqa: E731  (torch modules on a copy)

    y0 = net(x)
    assert (y0 - ref(net)).abs().max().item() < 1e-5
    twin = copy.deepcopy(net)                         # after the engine exists: the copy must not share it
    assert twin._eng is None and twin._flat_grad is None
    with torch.no_grad():
        for p in twin.parameters():
            p.mul_(1.1)
    assert (twin(x) - ref(twin)).abs().max().item() < 1e-5
    assert (net(x) - y0).abs().max().item() < 1e-6    # the original is unaffected by its twin's engine

    sd = {k: v.clone() for k, v in twin.state_dict().items()}
    net.load_state_dict(sd)                           # in-place copy into the bound tensors
    assert (net(x) - ref(twin)).abs().max().item() < 1e-5

    # a gradient tensor assigned from outside is adopted into the flat buffer and accumulated into
    for p in net.parameters():
        p.grad = None
    first = next(net.parameters())
    first.grad = torch.full_like(first, 0.25)
    y = net(x)
    y.sum().backward()
    want = copy.deepcopy(net)
    for p in want.parameters():
        p.grad = None
    nn.Sequential.forward(want, x).sum().backward()
    g_want = next(want.parameters()).grad
    assert torch.allclose(first.grad, g_want + 0.25, atol=1e-4 * max(1.0, g_want.abs().max().item()))
    assert first.grad.data_ptr() == net._grad_views[0].data_ptr()

    # a later forward replaces the kept pass: the older output's backward must say so instead of using the wrong activations
    ya = net(x)
    yb = net(x)
    with pytest.raises(RuntimeError, match="overwritten"):
        ya.sum().backward()
    yb.sum().backward()
    with torch.no_grad():
        yc = net(x)
    assert not yc.requires_grad and net._kept_token is None
