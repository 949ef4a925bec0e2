"""reset_net / set_step_mode / set_backend / detach_net — duck-typed exactly like spikingjelly's,
so the real package's functions and these are interchangeable on our modules."""
from torch import nn


def reset_net(net: nn.Module):
    for m in net.modules():
        if hasattr(m, "reset"):
            m.reset()


def set_step_mode(net: nn.Module, step_mode: str):
    for m in net.modules():
        if hasattr(m, "step_mode"):
            m.step_mode = step_mode


def set_backend(net: nn.Module, backend: str, instance=(nn.Module,)):
    for m in net.modules():
        if isinstance(m, instance) and hasattr(m, "backend") and backend in m.supported_backends:
            m.backend = backend


def detach_net(net: nn.Module):
    for m in net.modules():
        if hasattr(m, "detach"):
            m.detach()


def seq_to_ann_forward(x_seq, stateless_module):
    y = x_seq.flatten(0, 1)
    if isinstance(stateless_module, (list, tuple, nn.Sequential)):
        for m in stateless_module:
            y = m(y)
    else:
        y = stateless_module(y)
    return y.view(x_seq.shape[0], x_seq.shape[1], *y.shape[1:])
