"""Stand-in for spikingjelly.activation_based.layer.  TEST INFRASTRUCTURE ONLY.

Multi-step wrappers: in 'm' mode Conv2d / ConvTranspose2d / BatchNorm2d / GroupNorm take
[T, B, C, H, W], run the 4-D op on flatten(0, 1) and un-flatten (=> BN statistics over
T*B*H*W).  Linear is nn.Linear in either mode.  Dropout shares one mask over T.
"""
import torch
from torch import nn
import torch.nn.functional as F
from . import base, functional


class Linear(nn.Linear, base.StepModule):
    def __init__(self, in_features, out_features, bias=True, step_mode='s'):
        super().__init__(in_features, out_features, bias)
        self.step_mode = step_mode


def _multi_step(cls_name, torch_cls, ndim_m=5):
    class _Wrapped(torch_cls, base.StepModule):
        def __init__(self, *args, step_mode='s', **kwargs):
            super().__init__(*args, **kwargs)
            self.step_mode = step_mode

        def extra_repr(self):
            return super().extra_repr() + f', step_mode={self.step_mode}'

        def forward(self, x):
            if self.step_mode == 's':
                return super().forward(x)
            if x.dim() != ndim_m:
                raise ValueError(f'expected x with shape [T, N, C, H, W], but got x with shape {x.shape}!')
            return functional.seq_to_ann_forward(x, super().forward)

    _Wrapped.__name__ = cls_name
    _Wrapped.__qualname__ = cls_name
    return _Wrapped


Conv2d = _multi_step('Conv2d', nn.Conv2d)
ConvTranspose2d = _multi_step('ConvTranspose2d', nn.ConvTranspose2d)
BatchNorm2d = _multi_step('BatchNorm2d', nn.BatchNorm2d)


class GroupNorm(nn.GroupNorm, base.StepModule):
    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True, step_mode='s'):
        super().__init__(num_groups, num_channels, eps, affine)
        self.step_mode = step_mode

    def forward(self, x):
        if self.step_mode == 's':
            return super().forward(x)
        return functional.seq_to_ann_forward(x, super().forward)


class ThresholdDependentBatchNorm2d(BatchNorm2d):
    def __init__(self, alpha, v_th, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = alpha
        self.v_th = v_th
        assert self.affine
        torch.nn.init.constant_(self.weight, alpha * v_th)


class Dropout(base.MemoryModule):
    def __init__(self, p=0.5, step_mode='s'):
        super().__init__()
        self.step_mode = step_mode
        assert 0 <= p < 1
        self.register_memory('mask', None)
        self.p = p

    def extra_repr(self):
        return f'p={self.p}'

    def create_mask(self, x):
        self.mask = F.dropout(torch.ones_like(x.data), self.p, training=True)

    def single_step_forward(self, x):
        if self.training:
            if self.mask is None:
                self.create_mask(x)
            return x * self.mask
        return x

    def multi_step_forward(self, x_seq):
        if self.training:
            if self.mask is None:
                self.create_mask(x_seq[0])
            return x_seq * self.mask
        return x_seq
