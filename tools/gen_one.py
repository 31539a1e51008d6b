"""netG forward + backward of one call at a given size, three times (target of ncu captures of the generator kernels).
    python tools/gen_one.py <side>"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_image
from splice_b200.models.model import Model

side = int(sys.argv[1]) if len(sys.argv) > 1 else 224
torch.manual_seed(0)
net = Model(make_cfg("dino_vitb8")).netG
A = synth_image(1000, side, 8)[None].cuda()
for _ in range(3):
    out = net(A)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
