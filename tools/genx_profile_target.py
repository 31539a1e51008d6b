"""ncu target: inversion.py's generator (generalised native engine) at the script's real size, 224 x 298, batch 1: two eager
iterations (graphs off so that every kernel is its own launch) of forward + backward after two warm-up iterations.
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'convx|bn_bwd|cat_|wgrad|reflect_fold|sigmoid_bwd|update_running' \\
        --csv --log-file gpurun_out/genx_launches.csv python tools/genx_profile_target.py"""
import os
import sys
from pathlib import Path

os.environ.setdefault("SPLICE_B200_GRAPHS", "0")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

from splice_b200.inversion import NET_ARGS as INVERSION_ARGS  # noqa: E402
from splice_b200.models.unet.skip import skip  # noqa: E402

torch.manual_seed(0)
net = skip(32, 3, **INVERSION_ARGS).cuda()
x = torch.randn(1, 32, 224, 298, device="cuda")
target = torch.rand(1, 3, 224, 298, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for it in range(iters):
    for p in net.parameters():
        p.grad = None
    torch.nn.functional.mse_loss(net(x), target).backward()
torch.cuda.synchronize()
print("done", iters)
