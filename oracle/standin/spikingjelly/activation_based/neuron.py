"""Stand-in for spikingjelly.activation_based.neuron (0.0.0.0.14).  TEST INFRASTRUCTURE ONLY.

BaseNode: charge -> fire (surrogate(v - v_th)) -> reset (soft: v - s*v_th when v_reset is
None, hard: (1-s)*v + s*v_reset), reset path detached when ``detach_reset``.  Multi-step is
a python loop over x_seq[t] + torch.stack; ``v`` persists until ``reset()``.
LIF charge (decay_input): v + (x - v)/tau (v_reset None or 0) else v + (x - (v - v_reset))/tau.
IF charge: v + x.  PLIF: v + (x - v) * sigmoid(w), w = -log(init_tau - 1).
"""
import math
import torch
from torch import nn
from . import surrogate, base  # noqa: F401  (re-exported: reference imports them from here)


class BaseNode(base.MemoryModule):
    def __init__(self, v_threshold=1., v_reset=0., surrogate_function=surrogate.Sigmoid(),
                 detach_reset=False, step_mode='s', backend='torch', store_v_seq=False):
        assert isinstance(v_reset, float) or isinstance(v_reset, int) or v_reset is None
        assert isinstance(v_threshold, float) or isinstance(v_threshold, int)
        assert isinstance(detach_reset, bool)
        super().__init__()
        if v_reset is None:
            self.register_memory('v', 0.)
        else:
            self.register_memory('v', v_reset)
        self.v_threshold = v_threshold
        self.v_reset = v_reset
        self.detach_reset = detach_reset
        self.surrogate_function = surrogate_function
        self.step_mode = step_mode
        self.backend = backend
        self.store_v_seq = store_v_seq

    @property
    def store_v_seq(self):
        return self._store_v_seq

    @store_v_seq.setter
    def store_v_seq(self, value: bool):
        self._store_v_seq = value
        if value:
            if not hasattr(self, 'v_seq'):
                self.register_memory('v_seq', None)

    def neuronal_charge(self, x):
        raise NotImplementedError

    def neuronal_fire(self):
        return self.surrogate_function(self.v - self.v_threshold)

    def neuronal_reset(self, spike):
        spike_d = spike.detach() if self.detach_reset else spike
        if self.v_reset is None:
            self.v = self.v - spike_d * self.v_threshold
        else:
            self.v = (1. - spike_d) * self.v + spike_d * self.v_reset

    def extra_repr(self):
        return (f'v_threshold={self.v_threshold}, v_reset={self.v_reset}, detach_reset={self.detach_reset}, '
                f'step_mode={self.step_mode}, backend={self.backend}')

    def single_step_forward(self, x):
        self.v_float_to_tensor(x)
        self.neuronal_charge(x)
        spike = self.neuronal_fire()
        self.neuronal_reset(spike)
        return spike

    def multi_step_forward(self, x_seq):
        T = x_seq.shape[0]
        y_seq = []
        if self.store_v_seq:
            v_seq = []
        for t in range(T):
            y = self.single_step_forward(x_seq[t])
            y_seq.append(y)
            if self.store_v_seq:
                v_seq.append(self.v)
        if self.store_v_seq:
            self.v_seq = torch.stack(v_seq)
        return torch.stack(y_seq)

    def v_float_to_tensor(self, x):
        if isinstance(self.v, float) or isinstance(self.v, int):
            v_init = self.v
            self.v = torch.full_like(x.data, v_init)


class IFNode(BaseNode):
    @property
    def supported_backends(self):
        return ('torch', 'cupy')

    def neuronal_charge(self, x):
        self.v = self.v + x


class LIFNode(BaseNode):
    def __init__(self, tau=2., decay_input=True, v_threshold=1., v_reset=0.,
                 surrogate_function=surrogate.Sigmoid(), detach_reset=False, step_mode='s',
                 backend='torch', store_v_seq=False):
        assert isinstance(tau, float) and tau > 1.
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset, step_mode, backend, store_v_seq)
        self.tau = tau
        self.decay_input = decay_input

    @property
    def supported_backends(self):
        return ('torch', 'cupy')

    def extra_repr(self):
        return super().extra_repr() + f', tau={self.tau}'

    @staticmethod
    def neuronal_charge_decay_input_reset0(x, v, tau: float):
        return v + (x - v) / tau

    @staticmethod
    def neuronal_charge_decay_input(x, v, v_reset: float, tau: float):
        return v + (x - (v - v_reset)) / tau

    @staticmethod
    def neuronal_charge_no_decay_input_reset0(x, v, tau: float):
        return v * (1. - 1. / tau) + x

    @staticmethod
    def neuronal_charge_no_decay_input(x, v, v_reset: float, tau: float):
        return v - (v - v_reset) / tau + x

    def neuronal_charge(self, x):
        if self.decay_input:
            if self.v_reset is None or self.v_reset == 0.:
                self.v = self.neuronal_charge_decay_input_reset0(x, self.v, self.tau)
            else:
                self.v = self.neuronal_charge_decay_input(x, self.v, self.v_reset, self.tau)
        else:
            if self.v_reset is None or self.v_reset == 0.:
                self.v = self.neuronal_charge_no_decay_input_reset0(x, self.v, self.tau)
            else:
                self.v = self.neuronal_charge_no_decay_input(x, self.v, self.v_reset, self.tau)

    # eval-mode path of the real package: same arithmetic with spike = (v >= v_th)
    def _eval_step(self, x):
        self.neuronal_charge(x)
        spike = (self.v >= self.v_threshold).to(x)
        if self.v_reset is None:
            self.v = self.v - spike * self.v_threshold
        else:
            self.v = self.v_reset * spike + (1. - spike) * self.v
        return spike

    def single_step_forward(self, x):
        if self.training:
            return super().single_step_forward(x)
        self.v_float_to_tensor(x)
        return self._eval_step(x)

    def multi_step_forward(self, x_seq):
        if self.training:
            return super().multi_step_forward(x_seq)
        self.v_float_to_tensor(x_seq[0])
        spike_seq = torch.zeros_like(x_seq)
        v_seq = []
        for t in range(x_seq.shape[0]):
            spike_seq[t] = self._eval_step(x_seq[t])
            if self.store_v_seq:
                v_seq.append(self.v)
        if self.store_v_seq:
            self.v_seq = torch.stack(v_seq)
        return spike_seq


class ParametricLIFNode(BaseNode):
    def __init__(self, init_tau=2.0, decay_input=True, v_threshold=1., v_reset=0.,
                 surrogate_function=surrogate.Sigmoid(), detach_reset=False, step_mode='s',
                 backend='torch', store_v_seq=False):
        assert isinstance(init_tau, float) and init_tau > 1.
        super().__init__(v_threshold, v_reset, surrogate_function, detach_reset, step_mode, backend, store_v_seq)
        self.decay_input = decay_input
        init_w = - math.log(init_tau - 1.)
        self.w = nn.Parameter(torch.as_tensor(init_w))

    @property
    def supported_backends(self):
        return ('torch', 'cupy')

    def neuronal_charge(self, x):
        if self.decay_input:
            if self.v_reset is None or self.v_reset == 0.:
                self.v = self.v + (x - self.v) * self.w.sigmoid()
            else:
                self.v = self.v + (x - (self.v - self.v_reset)) * self.w.sigmoid()
        else:
            if self.v_reset is None or self.v_reset == 0.:
                self.v = self.v * (1. - self.w.sigmoid()) + x
            else:
                self.v = self.v - (self.v - self.v_reset) * self.w.sigmoid() + x
