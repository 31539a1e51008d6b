"""it/s of the feature-inversion loop on one GPU: splice_b200.inversion.invert (native generator_x + native ViT engine,
differentiable taps) against the same loop on stock PyTorch kernels (the reference's module-by-module generator + the oracle's DINO
ViT, fp32, torch defaults) - same image size (320x240 -> Resize(224) -> 224x298, t = 1037 for ViT-B/8), same feature, Adam lr 0.01.
Times iterations [warm, n) with CUDA events recorded once per iteration; stand-in ViT weights (offline).
    python tools/inversion_timing.py [dino_vitb8] [keys|cls] [n_iter]"""
import json
import sys
import tempfile
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from PIL import Image  # noqa: E402
from torchvision import transforms as T  # noqa: E402

from oracle import dino_vit, splice_ref as R  # noqa: E402
from oracle.make_golden_inversion import INVERSION_ARGS  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "dino_vitb8"
feature = sys.argv[2] if len(sys.argv) > 2 else "keys"
n_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 60
warm = 15
out = {"model": model, "feature": feature, "n_iter": n_iter, "timed_iters": n_iter - warm}

vsd = {k: v.detach() for k, v in dino_vit.build(model).cuda().state_dict().items()}
td = tempfile.mkdtemp()
rng = np.random.default_rng(0)
Image.fromarray(rng.integers(0, 256, (6, 8, 3), dtype=np.uint8)).resize((320, 240), Image.BICUBIC).save(f"{td}/in.png")


def it_per_s(events):
    torch.cuda.synchronize()
    return (len(events) - 1 - warm) / (events[warm].elapsed_time(events[-1]) * 1e-3)


# ---- product: splice_b200.inversion.invert ------------------------------------------------------------------
from splice_b200 import inversion  # noqa: E402

args = types.SimpleNamespace(feature=feature, layer=11, dino_model_name=model, image_path=f"{td}/in.png", save_path=f"{td}/out.png",
                             log_freq=10 ** 9, input_depth=32, LR=0.01, n_iter=n_iter, reduce_noise_stage_1_iter=10000,
                             reduce_noise_stage_2_iter=15000)
variants = [("native_it_s", {})]
if feature == "cls":   # the 'keys' mode adds no noise
    variants += [("native_it_s_inline_cpu_noise", {"prefetch_noise": False}), ("native_it_s_noise_on_device", {"noise_on_device": True})]
for key, kw in variants:
    events = []

    def cb(i, loss, net, net_input):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events.append(e)

    torch.manual_seed(0)
    _, losses = inversion.invert(args, vit_state_dict=vsd, callback=cb, **kw)
    out[key] = it_per_s(events)
    out[key.replace("it_s", "loss_first_last")] = [losses[0].item(), losses[-1].item()]

# ---- stock PyTorch: the reference's loop with torch modules + the oracle ViT ---------------------------------
from splice_b200.models.unet.skip import skip  # noqa: E402  (module tree only; evaluated by torch, not by the engine)

torch.manual_seed(0)
img = T.Compose([T.Resize(224), T.ToTensor()])(Image.open(f"{td}/in.png").convert("RGB")).unsqueeze(0).cuda()
net = skip(32, 3, **INVERSION_ARGS).cuda()
net_input_saved = torch.randn((1, 32, img.shape[-2], img.shape[-1])).cuda()
pre = T.Compose([T.Resize(224), T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
H = dino_vit.ARCH[model][2]


def feat(x):
    taps = R.vit_taps(vsd, pre(x))
    return taps["block"][11][:, 0, :] if feature == "cls" else R.keys_from_qkv(taps["qkv"][11], H)


with torch.no_grad():
    ref_feature = feat(img)
opt = torch.optim.Adam(net.parameters(), lr=0.01)
events = []
for i in range(n_iter):
    net_input = net_input_saved
    if feature == "cls":
        net_input = net_input_saved + (torch.randn(net_input_saved.shape).cuda() * 10)
    opt.zero_grad()
    loss = torch.nn.functional.mse_loss(feat(torch.nn.Sequential.forward(net, net_input)), ref_feature)
    loss.backward()
    opt.step()
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    events.append(e)
out["stock_pytorch_it_s"] = it_per_s(events)
out["speedup"] = out["native_it_s"] / out["stock_pytorch_it_s"]
print(json.dumps(out))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"inversion_timing_{model}_{feature}.json").write_text(json.dumps(out, indent=1))
