"""Per-segment timing of one optimisation step (diagnostic; run on the B200 box).
For each segment: host enqueue time (no sync) and device time (CUDA events, sync'd), averaged over steps."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_image, crop_schedule
from splice_b200.dino_init import random_dino_state_dict
from splice_b200.models.model import Model
from splice_b200.util.losses import LossG
from splice_b200.util.util import get_optimizer
from splice_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "dino_vitb8"
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
cfg = make_cfg(name)
torch.manual_seed(0)
model = Model(cfg)
crit = LossG(cfg, state_dict=random_dino_state_dict(name))
opt = get_optimizer(cfg, model.netG.parameters())
A, B = synth_image(1000, side, 8), synth_image(1001, side, 16)
sched = [(a.cuda(), b.cuda()) for a, b in crop_schedule(A, B, 16, 0)]
segs = ["zero_grad", "netG_fwd", "lossG", "backward", "adam"]
host = {s: 0.0 for s in segs}; dev = {s: 0.0 for s in segs}; launches = {s: 0 for s in segs}
def run(i, measure):
    a, b = sched[i % len(sched)]
    inputs = {"step": torch.tensor([float(i)]), "A_global": a, "B_global": b}
    st = {}
    def seg(nm, fn):
        if measure:
            torch.cuda.synchronize(); l0 = _lib.splice_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
        r = fn()
        if measure:
            e1.record(); host[nm] += time.perf_counter() - t0
            torch.cuda.synchronize(); dev[nm] += e0.elapsed_time(e1) * 1e-3; launches[nm] += _lib.splice_launch_count() - l0
        return r
    seg("zero_grad", lambda: opt.zero_grad())
    outs = seg("netG_fwd", lambda: model(inputs))
    losses = seg("lossG", lambda: crit(outs, inputs))
    seg("backward", lambda: losses["loss"].backward())
    seg("adam", lambda: opt.step())
for i in range(1, 30): run(i, False)
n = 50
for i in range(30, 30 + n):
    if i % 75 == 0: continue
    run(i, True)
print(f"{'segment':10s} {'host ms':>9s} {'device ms':>10s} {'launches':>9s}")
for s in segs:
    print(f"{s:10s} {host[s]/n*1e3:9.3f} {dev[s]/n*1e3:10.3f} {launches[s]/n:9.1f}")
print("total host", sum(host.values())/n*1e3, "device", sum(dev.values())/n*1e3)
