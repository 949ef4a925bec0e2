"""Surrogate-gradient descriptors (mirror of spikingjelly.activation_based.surrogate).

Inside the kernels only ``alpha`` and the kind are used; calling the object on a tensor (the
reference does so in the 'OR' connect function, Spiking_modules.py:868) runs heaviside forward
with the surrogate backward as a torch autograd function on the tensor's device.
"""
import math
import torch
from torch import nn
from .. import capi


def heaviside(x):
    return (x >= 0).to(x)


class _SG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha, kind):
        ctx.save_for_backward(x)
        ctx.alpha, ctx.kind = alpha, kind
        return heaviside(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        a = ctx.alpha
        if ctx.kind == capi.SDF_SG_ATAN:
            return a / 2 / (1 + (math.pi / 2 * a * x).pow(2)) * g, None, None
        s = (x * a).sigmoid()
        return g * (1.0 - s) * s * a, None, None


class SurrogateFunctionBase(nn.Module):
    kind = capi.SDF_SG_ATAN

    def __init__(self, alpha, spiking=True):
        super().__init__()
        self.alpha, self.spiking = alpha, spiking

    def set_spiking_mode(self, spiking):
        self.spiking = spiking

    def extra_repr(self):
        return f"alpha={self.alpha}, spiking={self.spiking}"

    def primitive_function(self, x):
        raise NotImplementedError

    def forward(self, x):
        if self.spiking:
            return _SG.apply(x, self.alpha, self.kind)
        return self.primitive_function(x)


class ATan(SurrogateFunctionBase):
    kind = capi.SDF_SG_ATAN

    def __init__(self, alpha=2.0, spiking=True):
        super().__init__(alpha, spiking)

    def primitive_function(self, x):
        return (math.pi / 2 * self.alpha * x).atan() / math.pi + 0.5


class Sigmoid(SurrogateFunctionBase):
    kind = capi.SDF_SG_SIGMOID

    def __init__(self, alpha=4.0, spiking=True):
        super().__init__(alpha, spiking)

    def primitive_function(self, x):
        return (x * self.alpha).sigmoid()
