"""Target of ncu runs: N device-resident optimisation steps of a bench.py workload (same objects and step function as
bench.py's timed loop, without priming / settling / the end-to-end and roofline legs, which take minutes under ncu).
    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 1600 --csv --log-file out.csv \
        python tools/profile_steps.py [--config 2] [--steps 45]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, crop_schedule, make_cfg, synth_pair  # noqa: E402
from splice_b200.dino_init import random_dino_state_dict  # noqa: E402
from splice_b200.models.model import Model  # noqa: E402
from splice_b200.util.losses import LossG  # noqa: E402
from splice_b200.util.util import get_optimizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--steps", type=int, default=45)
ap.add_argument("--sched", type=int, default=4, help="distinct crop pairs (each needs two steps before its graphs replay)")
args = ap.parse_args()
w = WORKLOADS[args.config]
cfg = make_cfg(w["model"])
cfg.update(dino_global_patch_size=w["vit_size"], global_A_crops_n_crops=w["n_crops"], global_B_crops_n_crops=w["n_crops"])
torch.manual_seed(0)
model = Model(cfg)
crit = LossG(cfg, state_dict=random_dino_state_dict(w["model"]))
opt = get_optimizer(cfg, model.netG.parameters())
A, B = synth_pair(w, 0)
sched = [(a.cuda(), b.cuda()) for a, b in crop_schedule(A, B, args.sched, seed=0, n_crops=w["n_crops"])]
torch.cuda.synchronize()
ready = torch.cuda.Event()
ready.record()
for a, b in sched:     # device-resident inputs, as in bench.py's resident leg
    a._splice_ready = b._splice_ready = ready
for i in range(1, 1 + args.steps):
    a, b = sched[i % len(sched)]
    inputs = {"step": torch.tensor([float(i)]), "A_global": a, "B_global": b}
    opt.zero_grad()
    losses = crit(model(inputs), inputs)
    losses["loss"].backward()
    opt.step()
torch.cuda.synchronize()
print("done", float(losses["loss"]))
