"""Multi-step wrappers of the stateless torch layers (mirror of spikingjelly layer.*).

These are parameter containers with the reference's state_dict names; the Swin blocks read
their tensors and run fused kernels instead of calling them.  Called directly they run the
cuDNN/cuBLAS op on flatten(0, 1) like spikingjelly (BN statistics over T*B*H*W)."""
import torch
from torch import nn
import torch.nn.functional as F
from . import base, functional


class Linear(nn.Linear, base.StepModule):
    def __init__(self, in_features, out_features, bias=True, step_mode="s"):
        super().__init__(in_features, out_features, bias)
        self.step_mode = step_mode


class _Seq5D(base.StepModule):
    def _fwd(self, x, op):
        if self.step_mode == "s":
            return op(x)
        if x.dim() != 5:
            raise ValueError(f"expected x with shape [T, N, C, H, W], but got x with shape {x.shape}!")
        return functional.seq_to_ann_forward(x, op)


def _plain(conv):
    return (conv.groups == 1 and tuple(conv.dilation) == (1, 1) and conv.padding_mode == "zeros"
            and not isinstance(conv.padding, str))


class Conv2d(nn.Conv2d, _Seq5D):
    """`spike_input = True` (set by the module that feeds it a neuron output) routes the call through
    ops.spike_conv2d: fp32-grade result from two TF32 tensor-core convolutions (exact for {0,1} inputs)."""
    spike_input = False

    def __init__(self, *a, step_mode="s", **k):
        super().__init__(*a, **k)
        self.step_mode = step_mode

    def _op(self, x):
        from .. import ops
        if self.spike_input and _plain(self):
            return ops.spike_conv2d(x, self.weight, self.bias, self.stride, self.padding)
        return ops.spike_conv2d(x, self.weight, self.bias, self.stride, self.padding, exact_input=False) \
            if _plain(self) else nn.Conv2d.forward(self, x)

    def forward(self, x):
        return self._fwd(x, self._op)


class ConvTranspose2d(nn.ConvTranspose2d, _Seq5D):
    spike_input = False

    def __init__(self, *a, step_mode="s", **k):
        super().__init__(*a, **k)
        self.step_mode = step_mode

    def _op(self, x):
        from .. import ops
        if _plain(self):
            return ops.spike_conv2d(x, self.weight, self.bias, self.stride, self.padding, True, self.output_padding,
                                    exact_input=self.spike_input)
        return nn.ConvTranspose2d.forward(self, x)

    def forward(self, x):
        return self._fwd(x, self._op)


class BatchNorm2d(nn.BatchNorm2d, _Seq5D):
    def __init__(self, *a, step_mode="s", **k):
        super().__init__(*a, **k)
        self.step_mode = step_mode

    def forward(self, x):
        return self._fwd(x, super().forward)


class GroupNorm(nn.GroupNorm, base.StepModule):
    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True, step_mode="s"):
        super().__init__(num_groups, num_channels, eps, affine)
        self.step_mode = step_mode

    def forward(self, x):
        if self.step_mode == "s":
            return super().forward(x)
        return functional.seq_to_ann_forward(x, super().forward)


class ThresholdDependentBatchNorm2d(BatchNorm2d):
    def __init__(self, alpha, v_th, *a, **k):
        super().__init__(*a, **k)
        self.alpha, self.v_th = alpha, v_th
        nn.init.constant_(self.weight, alpha * v_th)


class Dropout(base.MemoryModule):
    """One mask per reset, shared over T; identity in eval (and for p == 0)."""

    def __init__(self, p=0.5, step_mode="s"):
        super().__init__()
        assert 0 <= p < 1
        self.step_mode = step_mode
        self.register_memory("mask", None)
        self.p = p

    def extra_repr(self):
        return f"p={self.p}"

    def _masked(self, x, like):
        if not self.training or self.p == 0:
            return x
        if self.mask is None:
            self.mask = F.dropout(torch.ones_like(like.data), self.p, training=True)
        return x * self.mask

    def single_step_forward(self, x):
        return self._masked(x, x)

    def multi_step_forward(self, x_seq):
        return self._masked(x_seq, x_seq[0])
