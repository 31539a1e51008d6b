"""Device time of the generator alone (diagnostic; run on the B200 box): forward / backward of one call and of the two
parallel calls of a step, CUDA events around 50 repetitions."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import make_cfg, synth_image
from splice_b200.models.model import Model

cfg = make_cfg("dino_vitb8")
torch.manual_seed(0)
model = Model(cfg)
net = model.netG
SIDE = int(sys.argv[1]) if len(sys.argv) > 1 else 224
A = synth_image(1000, SIDE, 8)[None].cuda()
B = synth_image(1001, SIDE, 16)[None, :, :SIDE - 7, :SIDE - 7].contiguous().cuda()
print("generator input side", SIDE)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def fwd1():
    with torch.no_grad():
        return net(A)


def fwd2_keep():
    outs = net.forward_many([A, B])
    return outs


def fwdbwd2():
    outs = net.forward_many([A, B])
    torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])


def fwdbwd1():
    out = net(A)
    out.backward(torch.ones_like(out))


print("fwd 1 call (no grad) us:", round(timeit(fwd1), 1))
f2 = timeit(lambda: (fwd2_keep(), None)[1])
print("fwd 2 calls (kept, parallel streams) us:", round(f2, 1))
fb2 = timeit(fwdbwd2)
print("fwd+bwd 2 calls us:", round(fb2, 1), " => bwd ~", round(fb2 - f2, 1))
fb1 = timeit(fwdbwd1)
print("fwd+bwd 1 call us:", round(fb1, 1))
net.concurrent = False
print("fwd+bwd 2 calls sequential us:", round(timeit(fwdbwd2), 1))
