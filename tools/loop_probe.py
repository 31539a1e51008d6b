"""Diagnostic (B200 box): it/s of the bench's device-resident step loop at a given pair size, with the knobs that differ
between bench.py and tools/step_breakdown.py: crop-schedule length, targets' pass on a side stream, keys-only stop.
    python tools/loop_probe.py <side> [width]"""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_pair, crop_schedule
from splice_b200.dino_init import random_dino_state_dict
from splice_b200.models.model import Model
from splice_b200.util.losses import LossG
from splice_b200.util.util import get_optimizer

side = int(sys.argv[1]); width = int(sys.argv[2]) if len(sys.argv) > 2 else side
w = {"side": side, "width": width}
cfg = make_cfg("dino_vitb8")
torch.manual_seed(0)
model = Model(cfg)
crit = LossG(cfg, state_dict=random_dino_state_dict("dino_vitb8"))
opt = get_optimizer(cfg, model.netG.parameters())
A, B = synth_pair(w)
A_dev = A[None].cuda()


def run(nsched, overlap, n=150):
    crit.overlap_targets = overlap
    sched = [(a.cuda(), b.cuda()) for a, b in crop_schedule(A, B, nsched, seed=0)]
    def step(i):
        a, b = sched[i % len(sched)]
        inputs = {"step": torch.tensor([float(i)]), "A_global": a, "B_global": b}
        if i % 75 == 0:
            inputs["A"] = A_dev
        opt.zero_grad()
        losses = crit(model(inputs), inputs)
        losses["loss"].backward()
        opt.step()
    for i in range(1, 2 * nsched + 80):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(76, 76 + n):
        step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"side {side}x{width} sched {nsched} overlap {overlap}: {n / dt:.1f} it/s ({dt / n * 1e3:.2f} ms/step)", flush=True)


for nsched, overlap in ((16, True), (16, False), (32, True), (1, True), (1, False)):
    run(nsched, overlap)
