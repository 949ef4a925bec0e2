"""Stand-in for spikingjelly.activation_based.functional.  TEST INFRASTRUCTURE ONLY."""
import torch
from torch import nn
from . import base


def reset_net(net: nn.Module):
    for m in net.modules():
        if hasattr(m, 'reset'):
            m.reset()


def set_step_mode(net: nn.Module, step_mode: str):
    for m in net.modules():
        if hasattr(m, 'step_mode'):
            m.step_mode = step_mode


def set_backend(net: nn.Module, backend: str, instance=(nn.Module,)):
    for m in net.modules():
        if isinstance(m, instance):
            if hasattr(m, 'backend'):
                if backend in m.supported_backends:
                    m.backend = backend


def detach_net(net: nn.Module):
    for m in net.modules():
        if hasattr(m, 'detach'):
            m.detach()


def seq_to_ann_forward(x_seq: torch.Tensor, stateless_module):
    y_shape = [x_seq.shape[0], x_seq.shape[1]]
    y = x_seq.flatten(0, 1)
    if isinstance(stateless_module, (list, tuple, nn.Sequential)):
        for m in stateless_module:
            y = m(y)
    else:
        y = stateless_module(y)
    y_shape.extend(y.shape[1:])
    return y.view(y_shape)
