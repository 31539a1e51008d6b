"""compute-sanitizer target for the generalised generator engine (csrc/generator_x.cu): inversion.py's network forward + backward at
a small odd size, eagerly (graphs off) and twice (second pass accumulates), plus the small zero-padded batch-2 configuration.
    compute-sanitizer --tool memcheck python tools/genx_sanitize_target.py"""
import os
import sys
from pathlib import Path

os.environ.setdefault("SPLICE_B200_GRAPHS", "0")
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import torch  # noqa: E402

from splice_b200.inversion import NET_ARGS as INVERSION_ARGS  # noqa: E402
from splice_b200.models.unet.skip import skip  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"      # racecheck is slow: the inversion network only, one pass
torch.manual_seed(0)
for net, shape in ((skip(32, 3, **INVERSION_ARGS), (1, 32, 67, 90)),
                   (skip(5, 2, num_channels_down=[8, 12], num_channels_up=[8, 12], num_channels_skip=[3, 5], filter_size_down=[5, 3],
                         filter_size_up=[3, 7], filter_skip_size=3, need_sigmoid=False, pad="zero"), (2, 5, 45, 62))):
    net = net.cuda()
    x = torch.randn(*shape, device="cuda")
    for _ in range(1 if quick else 2):
        y = net(x)
        y.square().mean().backward()
    torch.cuda.synchronize()
    print("ok", shape, float(y.mean()), flush=True)
    if quick:
        break
