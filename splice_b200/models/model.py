"""Generator wrapper (drop-in for the reference's models/model.py:5-25)."""
from __future__ import annotations

import torch

from . import networks


class Model(torch.nn.Module):
    def __init__(self, cfg):
        super().__init__()
        device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.netG = networks.define_G(cfg['init_type'], cfg['init_gain']).to(device)
        self.cfg = cfg

    def forward(self, input):
        """{'x_global': netG(A_global), ['x_entire': netG(A)], 'y_global': netG(B_global)} (ref model.py:12-25)."""
        cfg = self.cfg
        if torch.cuda.is_available():
            # "inputs are on the device" marker: lets LossG start the targets' ViT pass on a side stream without waiting
            # for the generator work enqueued below (absent marker = it waits for the whole stream; same results)
            # (a caller that copies the inputs on its own stream attaches the event of that copy itself - train.py does -,
            # which lets the next step's targets start while this stream is still busy with the previous backward)
            ready = None
            for v in input.values():
                if torch.is_tensor(v) and v.is_cuda and getattr(v, "_splice_ready", None) is None:
                    if ready is None:
                        ready = torch.cuda.Event()
                        ready.record()
                    v._splice_ready = ready
        calls = []
        if cfg['lambda_global_cls'] + cfg['lambda_global_ssim'] > 0:
            calls.append(('x_global', input['A_global']))
        if cfg['lambda_entire_ssim'] > 0 and float(input['step']) % cfg['entire_A_every'] == 0:
            calls.append(('x_entire', input['A']))
        calls.append(('y_global', input['B_global']))
        # the 2-3 calls are independent (each has its own BatchNorm batch statistics): the native generator issues
        # them on parallel streams; running statistics are updated in call order like the reference's sequential calls
        many = getattr(self.netG, 'forward_many', None)
        outs = many([x for _, x in calls]) if many is not None else [self.netG(x) for _, x in calls]
        return {name: out for (name, _), out in zip(calls, outs)}
