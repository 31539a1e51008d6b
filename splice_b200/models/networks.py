"""Generator factory + weight initialisation (drop-in for the reference's models/networks.py:24-58).

Initialisation stays on torch's RNG stream, applied in `Module.apply` order over the same module tree, so a
given seed reproduces the reference's initial netG bit for bit (tests/test_generator_tree.py).
"""
from __future__ import annotations

from torch.nn import init

from .unet.skip import skip


def init_weights(net, init_type='normal', init_gain=0.02, debug=False):
    def init_func(m):
        classname = m.__class__.__name__
        is_conv_like = classname.find('Conv') != -1 or classname.find('Linear') != -1
        if hasattr(m, 'weight') and is_conv_like:
            if debug:
                print(classname)
            if init_type == 'normal':
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm2d') != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    net.apply(init_func)


def init_net(net, init_type='normal', init_gain=0.02, debug=False, initialize_weights=True):
    if initialize_weights:
        init_weights(net, init_type, init_gain=init_gain, debug=debug)
    return net


def define_G(init_type='normal', init_gain=0.02, initialize_weights=True):
    return init_net(skip(), init_type, init_gain, initialize_weights=initialize_weights)
