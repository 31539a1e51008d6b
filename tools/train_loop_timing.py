"""Wall-clock it/s of the complete drop-in loop `splice_b200.train.train_model` (dataset augmentation, input staging,
progress line, PNG + callback every `log_images_freq` steps) on a synthetic 224 px pair with DINO ViT-B/8 (seeded random
DINO-style weights): inline sampling like the reference vs the prefetching feed. Diagnostic; run on the B200 box.

    python tools/train_loop_timing.py [n]            # 224 px pair
    python tools/train_loop_timing.py fullres [n]    # 1200x900 pair, the reference's default regime (A_resize: -1): the PIL
                                                     # feed (prefetched) against the device-side feed (data/device_aug.py)"""
import os, sys, tempfile, time
from pathlib import Path
import numpy as np
import torch
from PIL import Image
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["SPLICE_B200_RANDOM_DINO"] = "1"
from splice_b200.train import train_model

argv = sys.argv[1:]
fullres = bool(argv) and argv[0] == "fullres"
if fullres:
    argv = argv[1:]
W, H = (1200, 900) if fullres else (224, 224)
root = Path(tempfile.mkdtemp())
for sub, seed, grid in (("A", 1000, 8), ("B", 1001, 16)):
    (root / sub).mkdir()
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (grid, grid, 3), dtype=np.uint8)
    img = np.asarray(Image.fromarray(low).resize((W, H), Image.BICUBIC)).astype(np.float64)
    Image.fromarray(np.clip(img + rng.normal(0, 8, img.shape), 0, 255).astype(np.uint8)).save(root / sub / "im.png")

n = int(argv[0]) if argv else (150 if fullres else 400)
if fullres:
    import yaml
    from splice_b200.data.Dataset import SingleImageDataset
    cfg = yaml.safe_load(open(ROOT / "splice_b200" / "conf" / "default" / "config.yaml"))
    cfg["dataroot"] = str(root)
    ds = SingleImageDataset(cfg)
    t0 = time.perf_counter()
    for _ in range(20):
        ds[0]
    print(f"RESULT PIL sample on the {W}x{H} pair: {(time.perf_counter() - t0) / 20 * 1e3:.1f} ms per sample (host, one thread)", flush=True)
    variants = (("PIL feed, prefetch=4", {}), ("device-side feed (device_aug), prefetch=4", {"device_aug": True}),
                ("device-side feed, image logging off", {"device_aug": True, "log_images_freq": 10 ** 9}))
else:
    variants = (("prefetch=4, async log (default)", {}), ("prefetch=4 in a forked worker", {"prefetch_mode": "process"}),
                ("prefetch=0 (inline sampling)", {"prefetch": 0}),
                ("prefetch=0, log_sync (reference loop semantics)", {"prefetch": 0, "log_sync": True}),
                ("prefetch=4, image logging off", {"log_images_freq": 10 ** 9}),
                ("device-side feed (device_aug), prefetch=4", {"device_aug": True}))
for label, ov in variants:
    ov = {"dino_model_name": "dino_vitb8", "n_epochs": n, "seed": 0, **ov}
    train_model(str(root), overrides={**ov, "n_epochs": 100 if fullres else 160})          # graph capture for every crop shape
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    train_model(str(root), overrides=ov)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"RESULT {label}: {n / dt:.1f} it/s ({dt / n * 1e3:.2f} ms/it incl. model construction)", flush=True)
