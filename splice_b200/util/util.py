"""Optimiser / scheduler factories and the PNG saver (drop-in for the reference's util/util.py:8-59)."""
from __future__ import annotations

from pathlib import Path

import torch
from torch.optim import lr_scheduler


def get_scheduler(optimizer, lr_policy, n_epochs=None, n_epochs_decay=None, lr_decay_iters=None):
    """ref util.py:8-25. The default policy 'none' is a constant LR via LambdaLR."""
    if lr_policy == 'linear':
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: max(1.0 - max(0, epoch) / float(n_epochs_decay + 1), 0))
    if lr_policy == 'step':
        return lr_scheduler.StepLR(optimizer, step_size=lr_decay_iters, gamma=0.5)
    if lr_policy == 'plateau':
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.2, threshold=0.01, patience=5)
    if lr_policy == 'cosine':
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=n_epochs, eta_min=0)
    if lr_policy == 'none':
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda x: 1)
    # the reference *returns* (does not raise) the exception object here (util.py:24) — behaviour kept
    return NotImplementedError('learning rate policy [%s] is not implemented', lr_policy)


def get_optimizer(cfg, params):
    """ref util.py:28-39. 'adam' is served by the fused multi-tensor sm_100a kernel (same update rule as
    torch.optim.Adam, same `param_groups` / `state_dict` layout); the other two stay torch's."""
    if cfg['optimizer'] == 'adam':
        from ..optim import FusedAdam

        return FusedAdam(params, lr=cfg['lr'], betas=(cfg['optimizer_beta1'], cfg['optimizer_beta2']))
    if cfg['optimizer'] == 'rmsprop':
        return torch.optim.RMSprop(params, lr=cfg['lr'])
    if cfg['optimizer'] == 'sgd':
        return torch.optim.SGD(params, lr=cfg['lr'])
    return NotImplementedError('optimizer [%s] is not implemented', cfg['optimizer'])


def save_result(image_t, dataroot):
    """Writes <dataroot>/out/output.png (ref util.py:55-59)."""
    from torchvision.transforms import ToPILImage

    out_dir = Path(f"{dataroot}/out")
    out_dir.mkdir(exist_ok=True, parents=True)
    ToPILImage()(image_t).save(f"{out_dir}/output.png")
