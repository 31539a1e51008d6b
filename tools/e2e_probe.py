"""Where does the end-to-end leg lose time against the device-resident leg? (diagnostic; run on the B200 box)
Variants of the step: inputs resident (with a completed ready event) / staged from pinned host memory each step on the
copy stream; loss read-back none / pinned ring of depth d / item()."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_image, crop_schedule
from splice_b200.dino_init import random_dino_state_dict
from splice_b200.models.model import Model
from splice_b200.util.losses import LossG
from splice_b200.util.util import get_optimizer, AsyncScalarLog, InputStager

cfg = make_cfg("dino_vitb8")
torch.manual_seed(0)
model = Model(cfg)
crit = LossG(cfg, state_dict=random_dino_state_dict("dino_vitb8"))
opt = get_optimizer(cfg, model.netG.parameters())
A, B = synth_image(1000, 224, 8), synth_image(1001, 224, 16)
host = [(a.pin_memory(), b.pin_memory()) for a, b in crop_schedule(A, B, 32, 0)]
dev = [(a.cuda(), b.cuda()) for a, b in host]
torch.cuda.synchronize()
ev = torch.cuda.Event(); ev.record()
for a, b in dev:
    a._splice_ready = ev; b._splice_ready = ev
stage = InputStager()


def step(i, copy, log):
    a, b = (host if copy else dev)[i % 32]
    inputs = {"step": torch.tensor([float(i)]), "A_global": a, "B_global": b}
    t0 = time.perf_counter()
    if copy:
        inputs = stage(inputs)
    opt.zero_grad()
    losses = crit(model(inputs), inputs)
    if log == "item":
        losses["loss"].item()
    elif log is not None:
        log.push(losses["loss"]); log.latest()
    losses["loss"].backward()
    opt.step()
    return time.perf_counter() - t0


def run(copy, log, n=200):
    for i in range(1, 40):
        if i % 75: step(i, copy, log)
    if isinstance(log, AsyncScalarLog): log.flush()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0; hostt = 0.0
    for i in range(40, 40 + n):
        if i % 75 == 0: continue
        hostt += step(i, copy, log); k += 1
    if isinstance(log, AsyncScalarLog): log.flush()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3, hostt / k * 1e3


for copy in (False, True):
    for name, log in (("none", None), ("ring2", AsyncScalarLog(2)), ("ring4", AsyncScalarLog(4)), ("ring8", AsyncScalarLog(8)), ("item", "item")):
        ms, h = run(copy, log)
        print(f"copy={copy!s:5} read={name:5}  {ms:.3f} ms/step   host {h:.3f} ms/step (incl. blocking)", flush=True)
