"""Where does the end-to-end leg lose time against the device-resident leg? (diagnostic; run on the B200 box)
Variants of the step: inputs resident / copied from pinned host memory each step, loss read-back none / async / item()."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_image, crop_schedule
from splice_b200.dino_init import random_dino_state_dict
from splice_b200.models.model import Model
from splice_b200.util.losses import LossG
from splice_b200.util.util import get_optimizer, AsyncScalarLog

cfg = make_cfg("dino_vitb8")
torch.manual_seed(0)
model = Model(cfg)
crit = LossG(cfg, state_dict=random_dino_state_dict("dino_vitb8"))
opt = get_optimizer(cfg, model.netG.parameters())
A, B = synth_image(1000, 224, 8), synth_image(1001, 224, 16)
host = [(a.pin_memory(), b.pin_memory()) for a, b in crop_schedule(A, B, 32, 0)]
dev = [(a.cuda(), b.cuda()) for a, b in host]
log = AsyncScalarLog()


def step(i, copy, read):
    a, b = (host if copy else dev)[i % 32]
    if copy:
        a, b = a.cuda(non_blocking=True), b.cuda(non_blocking=True)
    inputs = {"step": torch.tensor([float(i)]), "A_global": a, "B_global": b}
    opt.zero_grad()
    losses = crit(model(inputs), inputs)
    if read == "async":
        log.push(losses["loss"]); log.latest()
    elif read == "item":
        losses["loss"].item()
    losses["loss"].backward()
    opt.step()


def run(copy, read, n=150):
    for i in range(1, 40):
        if i % 75: step(i, copy, read)
    log.flush(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0
    for i in range(40, 40 + n):
        if i % 75 == 0: continue
        step(i, copy, read); k += 1
    log.flush(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3


for copy in (False, True):
    for read in ("none", "async", "item"):
        print(f"copy={copy!s:5} read={read:5}  {run(copy, read):.3f} ms/step", flush=True)
