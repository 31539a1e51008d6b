"""ORACLE tooling: validate oracle/splice_ref.py against the UNMODIFIED reference and write tests/golden/.

Runs only in the build container (needs /root/reference). `torch.hub.load` is resolved to the seeded DINO
stand-in of oracle/dino_vit.py because the hub repo/weights are unreachable offline.

    python -m oracle.make_golden            # validate + (re)write tests/golden/*.pt

What is pinned (all with the reference's own classes: VitExtractor, LossG, Model, define_G, get_optimizer):
  vit_s16.pt     ViT-S/16 taps on a seeded 224x224 input: keys self-sim, CLS token, layer-11 keys (summaries)
  resize.pt      global_transform on 3 input sizes (down-, up-scale, identity)
  step_s16.pt    one teacher-forced optimisation step at 128 px / ViT-S/16 for steps 0, 1 and 75:
                 inputs, per-term losses, dL/d(netG output) and netG gradient summaries, post-Adam parameters
"""
from __future__ import annotations

import contextlib
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))

from oracle import dino_vit, splice_ref as R  # noqa: E402


@contextlib.contextmanager
def reference_imported():
    """sys.path / cwd / torch.hub.load set up so that the reference imports and runs unmodified."""
    old_path, old_cwd, old_hub = list(sys.path), os.getcwd(), torch.hub.load
    sys.path.insert(0, str(REF))
    os.chdir(REF)
    torch.hub.load = dino_vit.hub_load_standin
    try:
        yield
    finally:
        torch.hub.load = old_hub
        os.chdir(old_cwd)
        sys.path[:] = old_path


def synth_image(seed: int, side: int, grid: int) -> torch.Tensor:
    """SURVEY.md §8d synthetic pair recipe, returned as a [3,side,side] float tensor in [0,1]."""
    from PIL import Image

    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (grid, grid, 3), dtype=np.uint8)
    img = np.asarray(Image.fromarray(low).resize((side, side), Image.BICUBIC)).astype(np.float64)
    img = np.clip(img + rng.normal(0, 8, img.shape), 0, 255).astype(np.uint8)
    return torch.from_numpy(img).permute(2, 0, 1).float() / 255.0


def summarize(t: torch.Tensor, n: int = 64) -> dict:
    """Compact fingerprint of a tensor: norm, sum, and n fixed pseudo-random samples."""
    flat = t.detach().reshape(-1).double()
    g = torch.Generator().manual_seed(flat.numel())
    idx = torch.randint(0, flat.numel(), (n,), generator=g)
    return {"shape": tuple(t.shape), "l2": flat.norm().item(), "sum": flat.sum().item(), "idx": idx,
            "samples": flat[idx].float()}


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main() -> None:
    torch.set_num_threads(8)
    GOLD.mkdir(parents=True, exist_ok=True)
    with reference_imported():
        from models.extractor import VitExtractor, attn_cosine_sim
        from models.model import Model
        from models.networks import define_G
        from util.losses import LossG
        from util.util import get_optimizer
        import yaml

        cfg = yaml.safe_load(open("conf/default/config.yaml"))

        # ---------------- ViT taps ----------------
        ext = VitExtractor("dino_vits16", "cpu")
        sd = {k: v.detach() for k, v in ext.model.state_dict().items()}
        img = R.global_transform(synth_image(7, 224, 8))[None]
        with torch.no_grad():
            ref_ssim = ext.get_keys_self_sim_from_input(img, 11)
            ref_keys = ext.get_keys_from_input(img, 11)
            ref_cls = ext.get_feature_from_input(img)[-1][0, 0, :]
            ref_attn = ext.get_attn_feature_from_input(img)[3]
            taps = R.vit_taps(sd, img)
            assert rel(R.keys_self_sim(sd, img), ref_ssim) < 1e-5
            assert rel(R.keys_from_qkv(taps["qkv"][11], 6), ref_keys) < 1e-5
            assert rel(taps["block"][-1][0, 0, :], ref_cls) < 1e-5
            assert rel(taps["attn"][3], ref_attn) < 1e-5
            assert rel(R.attn_cosine_sim(ref_keys.transpose(0, 1).reshape(1, 1, 197, 384)),
                       attn_cosine_sim(ref_keys.transpose(0, 1).reshape(1, 1, 197, 384))) < 1e-6
            # non-square input exercises the pos-embed interpolation
            img_ns = R.global_transform(synth_image(9, 224, 8))[None][:, :, :, :208]
            ref_ns = ext.get_keys_from_input(img_ns, 11)
            assert rel(R.keys_from_qkv(R.vit_taps(sd, img_ns)["qkv"][11], 6), ref_ns) < 1e-5
        torch.save({"img": img, "ssim": summarize(ref_ssim), "keys": summarize(ref_keys), "cls": ref_cls,
                    "attn3": summarize(ref_attn), "keys_nonsquare": summarize(ref_ns), "img_ns_cols": 208,
                    "ssim_full_row0": ref_ssim[0, 0].clone()}, GOLD / "vit_s16.pt")
        print("vit taps: restatement == reference")

        # ---------------- resize + normalise ----------------
        crit = LossG({**cfg, "dino_model_name": "dino_vits16"})
        rs = {}
        for side in (213, 224, 128, 448):
            x = synth_image(100 + side, side, 8)
            ref_t = crit.global_transform(x)
            got = R.global_transform(x)
            assert ref_t.shape == got.shape and (ref_t - got).abs().max().item() < 2e-5, (side, (ref_t - got).abs().max())
            rs[side] = {"inp": x if side <= 224 else None, "seed": 100 + side, "out": summarize(ref_t, 256)}
        xn = synth_image(55, 300, 8)[:, :225, :]  # non-square 225 x 300 -> 224 x 298
        ref_t = crit.global_transform(xn)
        got = R.global_transform(xn)
        assert ref_t.shape == got.shape == (3, 224, 298) and (ref_t - got).abs().max().item() < 2e-5
        rs["225x300"] = {"seed": 55, "out": summarize(ref_t, 256)}
        torch.save(rs, GOLD / "resize.pt")
        print("global_transform: restatement == reference")

        # ---------------- generator ----------------
        torch.manual_seed(0)
        netG = define_G(cfg["init_type"], cfg["init_gain"])
        gsd = {k: v.detach().clone() for k, v in netG.state_dict().items()}
        assert [k for k, _ in netG.named_parameters()] == R.generator_param_keys()
        for side in (128, 121):  # odd size exercises Concat's centre-crop
            x = synth_image(3, side, 8)[None]
            with torch.no_grad():
                ref_y = netG(x)
                got_y = R.generator_forward(gsd, x)
            assert ref_y.shape == got_y.shape and (ref_y - got_y).abs().max().item() < 1e-5, side
        print("generator: restatement == reference")

        # ---------------- teacher-forced steps ----------------
        cfg_s = {**cfg, "dino_model_name": "dino_vits16", "seed": 0}
        model = Model(cfg_s)
        model.netG.load_state_dict(gsd)
        crit = LossG(cfg_s)
        vsd = {k: v.detach() for k, v in crit.extractor.model.state_dict().items()}
        opt = get_optimizer(cfg_s, model.netG.parameters())
        A_full, B_full = synth_image(1000, 128, 8), synth_image(1001, 128, 16)
        steps = {}
        lam_state = None
        for step, (sa, sb) in {0: (128, 125), 1: (123, 128), 75: (126, 122)}.items():
            model.netG.load_state_dict(gsd)  # teacher forcing: identical parameters at every pinned step
            inputs = {"step": torch.tensor([float(step)]), "A_global": A_full[None, :, :sa, :sa].contiguous(),
                      "B_global": B_full[None, :, :sb, :sb].contiguous()}
            if step % cfg_s["entire_A_every"] == 0:
                inputs["A"] = A_full[None]
            opt.zero_grad()
            outputs = model(inputs)
            for v in outputs.values():
                v.retain_grad()
            losses = crit(outputs, inputs)
            losses["loss"].backward()
            # restatement on the same inputs
            lam_state = R.active_lambdas(cfg_s, step, lam_state)
            assert {k: float(v) for k, v in lam_state.items()} == {k: float(v) for k, v in crit.lambdas.items()}
            o_outputs = {k: R.generator_forward(gsd, inputs[{"x_global": "A_global", "y_global": "B_global",
                                                             "x_entire": "A"}[k]]).requires_grad_(True)
                         for k in outputs}
            o_losses = R.loss_g(vsd, cfg_s, lam_state, o_outputs, inputs)
            o_losses["loss"].backward()
            for k in losses:
                assert abs(float(losses[k].detach()) - float(o_losses[k].detach())) <= 1e-5 * max(1.0, abs(float(losses[k].detach()))), (step, k)
            for k in outputs:
                if outputs[k].grad is None:  # e.g. y_global at step 0: the identity term is still off
                    assert o_outputs[k].grad is None
                    continue
                assert rel(o_outputs[k].grad, outputs[k].grad) < 1e-3, (step, k, rel(o_outputs[k].grad, outputs[k].grad))
            rec = {"inputs": {k: v.clone() for k, v in inputs.items()},
                   "losses": {k: float(v) for k, v in losses.items()},
                   "out": {k: summarize(v, 256) for k, v in outputs.items()},
                   "dout": {k: summarize(v.grad, 256) for k, v in outputs.items() if v.grad is not None},
                   "dout_x_global": outputs["x_global"].grad.clone(),
                   "grads": {k: summarize(p.grad) for k, p in model.netG.named_parameters()}}
            if step == 1:
                # one Adam step from fresh optimiser state (step count 1) — pins util.py:30-32 / adam.py
                opt2 = get_optimizer(cfg_s, model.netG.parameters())
                grads = {k: p.grad.detach().clone() for k, p in model.netG.named_parameters()}
                opt2.step()
                for k, p in model.netG.named_parameters():
                    q, m, v = gsd[k].clone(), torch.zeros_like(p), torch.zeros_like(p)
                    R.adam_step(q, grads[k], m, v, 1, cfg_s["lr"], cfg_s["optimizer_beta1"], cfg_s["optimizer_beta2"])
                    assert (q - p.detach()).abs().max().item() < 1e-7, k
                rec["post_adam"] = {k: summarize(p) for k, p in model.netG.named_parameters()}
                rec["grad_full_small"] = {k: grads[k] for k in ("9.0.weight", "9.0.bias", "1.0.1.0.weight", "6.0.weight")}
            steps[step] = rec
            print(f"step {step}: losses", {k: round(float(v), 6) for k, v in losses.items()})
        torch.save({"cfg": cfg_s, "netG": gsd, "steps": steps}, GOLD / "step_s16.pt")
        print("teacher-forced steps: restatement == reference")
    for f in sorted(GOLD.glob("*.pt")):
        print(f.name, f.stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
