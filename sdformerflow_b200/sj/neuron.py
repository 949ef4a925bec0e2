"""Spiking neurons backed by sdf_lif_fwd / sdf_lif_bwd (mirror of spikingjelly neuron.*).

Same constructor signatures, attributes (v_threshold, v_reset, tau, detach_reset,
surrogate_function, step_mode, backend, store_v_seq) and memory protocol as spikingjelly's
BaseNode / IFNode / LIFNode / ParametricLIFNode; the charge-fire-reset loop over T runs in one
CUDA kernel with the membrane potential in registers.  CPU tensors raise (no fallback).
"""
import math
import torch
from torch import nn
from . import surrogate, base  # noqa: F401  (the reference imports them from this module)
from .. import capi, ops


class BaseNode(base.MemoryModule):
    kind = None

    def __init__(self, v_threshold=1.0, v_reset=0.0, surrogate_function=None, detach_reset=False, step_mode="s",
                 backend="torch", store_v_seq=False):
        assert isinstance(v_reset, (float, int)) or v_reset is None
        assert isinstance(v_threshold, (float, int))
        assert isinstance(detach_reset, bool)
        super().__init__()
        self.register_memory("v", 0.0 if v_reset is None else v_reset)
        self.v_threshold, self.v_reset, self.detach_reset = v_threshold, v_reset, detach_reset
        self.surrogate_function = surrogate_function if surrogate_function is not None else surrogate.Sigmoid()
        self.step_mode, self.backend = step_mode, backend
        self.store_v_seq = store_v_seq

    @property
    def store_v_seq(self):
        return self._store_v_seq

    @store_v_seq.setter
    def store_v_seq(self, value):
        self._store_v_seq = value
        if value and not hasattr(self, "v_seq"):
            self.register_memory("v_seq", None)

    def extra_repr(self):
        return (f"v_threshold={self.v_threshold}, v_reset={self.v_reset}, detach_reset={self.detach_reset}, "
                f"step_mode={self.step_mode}, backend={self.backend}")

    def _tau(self):
        return 2.0

    def _plif_w(self):
        return None

    def neuron_cfg(self):
        sf = self.surrogate_function
        return ops.NeuronCfg(kind=self.kind, v_th=float(self.v_threshold), v_reset=self.v_reset, tau=self._tau(),
                             detach_reset=self.detach_reset, surrogate=getattr(sf, "kind", capi.SDF_SG_ATAN),
                             sg_alpha=float(getattr(sf, "alpha", 2.0)))

    def multi_step_forward(self, x_seq):
        v_init = self.v.reshape(-1).contiguous() if isinstance(self.v, torch.Tensor) else None
        spike, v = ops.neuron(x_seq, self.neuron_cfg(), 0, self._plif_w(), v_init, want_state=True)
        self.v = v.view(x_seq.shape[1:])
        if self.store_v_seq:
            raise NotImplementedError("store_v_seq: use sdformerflow_b200.ops.neuron_debug for membrane traces")
        return spike

    def single_step_forward(self, x):
        return self.multi_step_forward(x.unsqueeze(0))[0]


class IFNode(BaseNode):
    kind = capi.SDF_NEURON_IF


class LIFNode(BaseNode):
    kind = capi.SDF_NEURON_LIF

    def __init__(self, tau=2.0, decay_input=True, v_threshold=1.0, v_reset=0.0, surrogate_function=None,
                 detach_reset=False, step_mode="s", backend="torch", store_v_seq=False):
        assert isinstance(tau, float) and tau > 1.0
        if not decay_input:
            raise NotImplementedError("LIFNode(decay_input=False) is not used by SDformerFlow and not built")
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset, step_mode, backend, store_v_seq)
        self.tau, self.decay_input = tau, decay_input

    def _tau(self):
        return self.tau

    def extra_repr(self):
        return super().extra_repr() + f", tau={self.tau}"


class ParametricLIFNode(BaseNode):
    kind = capi.SDF_NEURON_PLIF

    def __init__(self, init_tau=2.0, decay_input=True, v_threshold=1.0, v_reset=0.0, surrogate_function=None,
                 detach_reset=False, step_mode="s", backend="torch", store_v_seq=False):
        assert isinstance(init_tau, float) and init_tau > 1.0
        if not decay_input:
            raise NotImplementedError("ParametricLIFNode(decay_input=False) is not built")
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset, step_mode, backend, store_v_seq)
        self.decay_input = decay_input
        self.w = nn.Parameter(torch.as_tensor(-math.log(init_tau - 1.0)))

    def _plif_w(self):
        return self.w

    def _tau(self):
        # the kernels take 1/k = 1/sigmoid(w) by value (one host read per parameter version)
        return ops.plif_tau(self.w) if self.w.is_cuda else 1.0 / torch.sigmoid(self.w.detach()).item()
