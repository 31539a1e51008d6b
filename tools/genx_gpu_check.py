"""The parts of tools/gpu_checks.py: generator_inversion_variant (inversion.py's generator on the generalised native
engine, csrc/generator_x.cu). Each part appends rows {"part", "ok", ...}; a part that raises is recorded and the others still run."""
from __future__ import annotations

import copy
import sys
import tempfile
import traceback
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _maxabs(a, b):
    return (a.float() - b.float()).abs().max().item()


def part_golden(rows):
    """(a) against the golden the UNMODIFIED reference produced (tests/golden/inversion_gen.pt)."""
    import torch
    from oracle.make_golden_inversion import INVERSION_ARGS, golden_input, perturb
    from splice_b200.generator_x import NativeSkipX
    from splice_b200.models.unet.skip import skip

    gold = torch.load(ROOT / "tests" / "golden" / "inversion_gen.pt", weights_only=False)
    torch.manual_seed(0)
    net = skip(32, 3, **INVERSION_ARGS)
    assert isinstance(net, NativeSkipX) and list(net.state_dict().keys()) == gold["keys"]
    perturb(net, 1)
    net = net.cuda()
    x, w = (t.cuda() for t in golden_input())
    y = net(x)
    (y * w).sum().backward()
    numel = {k: p.numel() for k, p in net.named_parameters()}
    big = max(fp["abs"] / numel[k] for k, fp in gold["grads"].items())

    def close(fp, t, rel):
        f = t.detach().reshape(-1).double().cpu()
        idx = torch.linspace(0, f.numel() - 1, min(16, f.numel())).long()
        tol = rel * max(fp["abs"] / max(f.numel(), 1), 1e-6)
        return abs(f.sum().item() - fp["sum"]) <= rel * max(fp["abs"], 1e-6) and (f[idx] - fp["samples"]).abs().max().item() <= 50 * tol

    bad = []
    for k, p in net.named_parameters():
        fp = gold["grads"][k]
        # conv biases in front of a BatchNorm have an exactly-zero gradient (both sides return cancellation noise for them)
        ok = (p.grad.abs().mean().item() < 1e-3 * big) if fp["abs"] / numel[k] < 1e-4 * big else close(fp, p.grad, 5e-2)
        if not ok:
            bad.append(k)
    bad_buf = [k for k, v in net.state_dict().items() if k in gold["buffers"] and not close(gold["buffers"][k], v.float(), 1e-3)]
    r = {"part": "golden", "out_maxabs": _maxabs(y.cpu(), gold["y"]), "bad_grads": bad[:5], "n_bad_grads": len(bad),
         "bad_buffers": bad_buf[:5]}
    r["ok"] = r["out_maxabs"] < 5e-4 and not bad and not bad_buf
    rows.append(r)


def part_fp64(rows, full=(224, 298)):
    """(b) against torch evaluating the same module tree in float64, torch's own float32 (cuDNN, TF32 off) as the yardstick."""
    import torch
    from oracle.make_golden_inversion import INVERSION_ARGS
    from splice_b200.models.unet.skip import skip
    from tools.genx_compare import compare, randomise

    for (h, wd) in ((67, 90), full):
        torch.manual_seed(0)
        net = skip(32, 3, **INVERSION_ARGS)
        randomise(net, 1)
        net = net.cuda()
        xx = torch.randn(1, 32, h, wd, generator=torch.Generator().manual_seed(2)).cuda()
        try:
            r = compare(net, xx, 3)
            for p in net.parameters():
                p.grad = None
            r2 = compare(net, xx + 0.5 * torch.randn(xx.shape, generator=torch.Generator().manual_seed(4)).cuda(), 5)
            r.update(part="fp64", H=h, W=wd, second_pass_grad_l2_rel=r2["grad_l2_rel"], ok=True)
        except AssertionError as e:
            r = {"part": "fp64", "H": h, "W": wd, "ok": False, "error": str(e)[:300]}
        rows.append(r)
    # a small zero-padded, batch-2, mixed-filter configuration (the other branches of the generalised kernels)
    torch.manual_seed(10)
    net = skip(5, 2, num_channels_down=[8, 12], num_channels_up=[8, 12], num_channels_skip=[3, 5], filter_size_down=[5, 3],
               filter_size_up=[3, 7], filter_skip_size=3, need_sigmoid=False, pad="zero")
    randomise(net, 11)
    net = net.cuda()
    xx = torch.randn(2, 5, 45, 62, generator=torch.Generator().manual_seed(12)).cuda()
    try:
        r = compare(net, xx, 13)
        r2 = compare(net, xx, 14)      # gradients left in place: accumulated into
        r.update(part="fp64_zero_pad_batch2", second_pass_grad_l2_rel=r2["grad_l2_rel"], ok=True)
    except AssertionError as e:
        r = {"part": "fp64_zero_pad_batch2", "ok": False, "error": str(e)[:300]}
    rows.append(r)


def part_adam_loop(rows, full=(224, 298), n_time=30):
    """(c) six Adam iterations of the generator alone (graph capture on the 2nd, replay from the 3rd) against the torch-module copy
    of the same loop, then 30 timed iterations of each (CUDA events)."""
    import torch
    from oracle.make_golden_inversion import INVERSION_ARGS
    from splice_b200.models.unet.skip import skip

    torch.manual_seed(0)
    net = skip(32, 3, **INVERSION_ARGS).cuda()
    ref = copy.deepcopy(net)
    g = torch.Generator().manual_seed(7)
    z = torch.randn(1, 32, *full, generator=g).cuda()
    target = torch.rand(1, 3, *full, generator=g).cuda()
    noises = [torch.randn(z.shape, generator=g).cuda() for _ in range(6)]

    def loop(model, forward, n_time):
        opt = torch.optim.Adam(model.parameters(), lr=0.01)

        def step(it):
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(forward(model, z + 2 * noises[it % 6]), target)
            loss.backward()
            opt.step()
            return loss

        out = [step(it).item() for it in range(6)]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(n_time):
            step(it)
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / n_time

    l_native, ms_native = loop(net, lambda m, a: m(a), n_time)
    l_ref, ms_ref = loop(ref, lambda m, a: torch.nn.Sequential.forward(m, a), n_time)
    rel = [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(l_native, l_ref)]
    r = {"part": "adam_loop", "loss_native": l_native, "loss_torch_modules": l_ref, "loss_rel": rel,
         "ms_per_iter_native": ms_native, "ms_per_iter_torch_modules_fp32": ms_ref}
    r["ok"] = rel[0] < 1e-4 and max(rel) < 5e-2 and all(v == v for v in l_native)
    rows.append(r)


def part_invert(rows):
    """(d) splice_b200.inversion.invert() end to end for both feature kinds on a synthetic image (stand-in ViT-S/16 weights)."""
    import numpy as np
    import torch
    from PIL import Image

    from oracle import dino_vit
    from splice_b200 import inversion

    vsd = {k: v.detach() for k, v in dino_vit.build("dino_vits16").cuda().state_dict().items()}
    with tempfile.TemporaryDirectory() as td:
        rng = np.random.default_rng(0)
        low = rng.integers(0, 256, (6, 8, 3), dtype=np.uint8)
        Image.fromarray(low).resize((320, 240), Image.BICUBIC).save(f"{td}/in.png")
        for feature in ("keys", "cls"):
            args = types.SimpleNamespace(feature=feature, layer=11, dino_model_name="dino_vits16", image_path=f"{td}/in.png",
                                         save_path=f"{td}/out_{feature}.png", log_freq=4, input_depth=32, LR=0.01, n_iter=8,
                                         reduce_noise_stage_1_iter=3, reduce_noise_stage_2_iter=6)
            torch.manual_seed(1)
            _, losses = inversion.invert(args, vit_state_dict=vsd)
            out_img = Image.open(args.save_path)
            r = {"part": "invert_" + feature, "losses": losses.tolist(), "saved_size": list(out_img.size)}
            r["ok"] = bool(torch.isfinite(losses).all()) and len(losses) == 8 and tuple(out_img.size) == (298, 224)
            rows.append(r)


def run_all():
    import torch

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rows = []
    for fn in (part_golden, part_fp64, part_adam_loop, part_invert):
        try:
            fn(rows)
        except Exception as e:  # noqa: BLE001
            rows.append({"part": fn.__name__, "ok": False, "error": f"{type(e).__name__}: {e}"[:400],
                         "trace": traceback.format_exc()[-1500:]})
    return rows


if __name__ == "__main__":
    import json

    print(json.dumps(run_all(), indent=1))
