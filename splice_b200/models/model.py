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
        outputs = {}
        if cfg['lambda_global_cls'] + cfg['lambda_global_ssim'] > 0:
            outputs['x_global'] = self.netG(input['A_global'])
        if cfg['lambda_entire_ssim'] > 0 and float(input['step']) % cfg['entire_A_every'] == 0:
            outputs['x_entire'] = self.netG(input['A'])
        outputs['y_global'] = self.netG(input['B_global'])
        return outputs
